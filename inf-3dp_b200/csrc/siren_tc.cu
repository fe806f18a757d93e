// tcgen05 tensor-core engine of the fused SIREN jet -- placeholder until the kernel lands (next commit).
#include "common.cuh"

namespace inf3dp {

size_t siren_tc_section_bytes() { return 0; }

int siren_tc_pack(const float* const*, const float* const*, float, int, void*, const SirenLayout&, SirenHeader& hdr,
                  cudaStream_t) {
    hdr.tc_ok = 0;
    return INF3DP_OK;
}

int siren_tc_launch(const void*, const SirenHeader&, const float*, int64_t, int, float*, float*, bool, cudaStream_t) {
    set_error("tensor-core engine not built");
    return INF3DP_EINVAL;
}

}  // namespace inf3dp
