// Shared host/device helpers of libinf3dp (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/inf3dp.h"

namespace inf3dp {

void set_error(const char* fmt, ...);

#define INF3DP_CHECK_ARG(cond, ...)                      \
    do {                                                 \
        if (!(cond)) {                                   \
            ::inf3dp::set_error(__VA_ARGS__);            \
            return INF3DP_EINVAL;                        \
        }                                                \
    } while (0)

#define INF3DP_CUDA(call)                                                                        \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            ::inf3dp::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return INF3DP_ECUDA;                                                                 \
        }                                                                                        \
    } while (0)

#define INF3DP_LAUNCH_CHECK(name)                                                                \
    do {                                                                                         \
        cudaError_t e__ = cudaGetLastError();                                                    \
        if (e__ != cudaSuccess) {                                                                \
            ::inf3dp::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));       \
            return INF3DP_ECUDA;                                                                 \
        }                                                                                        \
    } while (0)

inline int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

// SMs the persistent SIREN engines may occupy (inf3dp_siren_set_sm_limit; 0 = all).  A multi-GPU caller leaves a few SMs
// to the NCCL kernels of an all-gather that overlaps the next chunk's query.
int siren_sm_limit();   // api.cu
inline int siren_sms() {
    const int all = num_sms(), lim = siren_sm_limit();
    return (lim > 0 && lim < all) ? lim : all;
}

constexpr size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------------------
// Packed SIREN weights (built by inf3dp_siren_pack_weights, consumed by every engine).
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t kSirenMagic = 0x33445031u;  // "1PD3"
constexpr int kHidden = 256;
constexpr int kMaxOut = 64;
constexpr int kMaxHiddenLayers = 16;

struct SirenHeader {
    uint32_t magic;
    int in_features, hidden, num_hidden_layers, out_features, activation;
    float omega0;
    int tc_ok;             // flagship configuration: the tcgen05 engine may be used
    uint64_t off_w0;       // f32 [hidden][4]   first layer, omega folded (sine), input padded to 4
    uint64_t off_b0;       // f32 [hidden]
    uint64_t off_wt;       // f32 [L][hidden k][hidden n]   hidden layers, transposed, omega folded
    uint64_t off_b;        // f32 [L][hidden]
    uint64_t off_wout;     // f32 [kMaxOut][hidden]  rows >= out_features are zero
    uint64_t off_bout;     // f32 [kMaxOut]
    uint64_t off_tc;       // tcgen05 operand images (see siren_tc.cu), 0 if !tc_ok
    uint64_t total_bytes;
    float tc_scale;        // power-of-two pre-scale applied to the FP16 hidden weights (undone in the epilogue)
    float tc_inv_scale;
    uint64_t off_rev;      // reverse-mode tcgen05 operand images (see siren_rev.cu), 0 if !rev_ok
    int rev_ok;            // flagship configuration with out_features == 1: the reverse-mode engine may be used
    int mlp_ok;            // ReLU classifier 4 -> 256 -> 256 -> C (C <= 16): off_rev holds its CTA-pair operand image
    uint32_t pad[24];
};
static_assert(sizeof(SirenHeader) <= 256, "header must fit 256 bytes");

struct SirenLayout {
    size_t off_w0, off_b0, off_wt, off_b, off_wout, off_bout, off_tc, off_rev, total;
    bool tc_ok, rev_ok, mlp_ok;
};

size_t siren_tc_section_bytes();   // siren_tc.cu
size_t siren_rev_section_bytes();  // siren_rev.cu
size_t siren_rev_workspace_bytes();
size_t siren_mlp_section_bytes();  // siren_rev.cu

inline bool siren_config_supported(int in_features, int hidden, int L, int out) {
    return hidden == kHidden && (in_features == 3 || in_features == 4) && out >= 1 && out <= kMaxOut && L >= 0 &&
           L <= kMaxHiddenLayers;
}

inline SirenLayout siren_layout(int in_features, int hidden, int L, int out, int activation) {
    SirenLayout l{};
    size_t o = 256;
    l.off_w0 = o;   o += (size_t)hidden * 4 * 4;
    l.off_b0 = o;   o += (size_t)hidden * 4;
    l.off_wt = o;   o += (size_t)L * hidden * hidden * 4;
    l.off_b = o;    o += (size_t)(L > 0 ? L : 1) * hidden * 4;
    l.off_wout = o; o += (size_t)kMaxOut * hidden * 4;
    l.off_bout = o; o += (size_t)kMaxOut * 4;
    o = align_up(o, 1024);
    l.tc_ok = (in_features == 3 && hidden == kHidden && L == 3 && out <= 4 && activation == INF3DP_ACT_SINE);
    l.off_tc = o;
    // the TC section is always reserved for in=3/L=3 so the blob size does not depend on the activation
    if (in_features == 3 && hidden == kHidden && L == 3 && out <= 8) o += siren_tc_section_bytes();
    o = align_up(o, 1024);
    l.rev_ok = l.tc_ok && out == 1;
    l.off_rev = o;
    if (in_features == 3 && hidden == kHidden && L == 3 && out == 1) o += siren_rev_section_bytes();
    // level classifier (exp_scripts/train_levels.py:35-51): reserved independently of the activation, like the TC section
    l.mlp_ok = (in_features == 4 && hidden == kHidden && L == 1 && out <= 16 && activation == INF3DP_ACT_RELU);
    if (in_features == 4 && hidden == kHidden && L == 1 && out <= 16) o += siren_mlp_section_bytes();
    l.total = align_up(o, 256);
    return l;
}

// engine entry points (implemented in the respective .cu files)
int siren_simt_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, int order, float* y,
                      float* J, float* H, cudaStream_t stream);
int siren_tc_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, int order, float* y,
                    float* J, bool precise, cudaStream_t stream);
int siren_tc_pack(const float* const* W, const float* const* b, float omega0, int out_features, void* packed,
                  const SirenLayout& lay, SirenHeader& hdr, cudaStream_t stream);
const uint32_t* siren_tc_maxabs_bits(const void* packed, const SirenLayout& lay);  // device uint32[4], valid after siren_tc_pack
int siren_rev_pack(const float* const* W, float omega0, const uint32_t* maxabs_bits, void* packed, const SirenLayout& lay,
                   SirenHeader& hdr, cudaStream_t stream);
int siren_rev_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, float* J, void* ws,
                     size_t ws_bytes, cudaStream_t stream);
int siren_rev_launch_rows(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* rows, void* ws,
                          size_t ws_bytes, cudaStream_t stream);
int siren_rev_launch_rows_peers(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* const* peer_rows,
                                int n_peers, float* multicast_rows, void* ws, size_t ws_bytes, cudaStream_t stream);
int siren_mlp_pack(const float* const* W, const float* const* b, int out_features, void* packed, const SirenLayout& lay,
                   SirenHeader& hdr, cudaStream_t stream);
int siren_mlp_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, cudaStream_t stream);
int siren_jet2_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, float* J, float* H,
                      cudaStream_t stream);
int siren_value_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, cudaStream_t stream);
int siren_hess_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, float* J, float* H,
                      void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace inf3dp
