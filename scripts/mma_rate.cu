// Developer microbenchmark: issue rate of tcgen05.mma kind::f16 in the operand configurations the SIREN engines use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I inf-3dp_b200/csrc scripts/mma_rate.cu -o /tmp/mma_rate
// Prints cycles per MMA (M = 128 rows per CTA, K = 16) for: cta_group 1 / 2, A from TMEM / SMEM, N = 256 / 128,
// one accumulator / two alternating accumulators.  Operand contents are irrelevant (zeros).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "tc_ptx.cuh"

using namespace inf3dp::tcptx;

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss_pair(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// variant bits: 1 = A from smem (else TMEM), 2 = N=128 (else 256), 4 = two alternating accumulators, 8 = same A/B every MMA
template <int CTAS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int variant, int n_mma, unsigned long long* out) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* sptr = smem_dyn + (sbase - smem_u32(smem_dyn));
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sptr)[i] = 0u;
    const uint32_t bar = sbase + 200 * 1024, tptr = sbase + 200 * 1024 + 64;
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (CTAS == 2) cluster_sync_all();
    if (warp == 0) {
        if (CTAS == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sptr + 200 * 1024 + 64);
    const bool a_smem = variant & 1, n128 = variant & 2, two_acc = variant & 4, same = variant & 8;
    const int N = n128 ? 128 : 256;
    const int nb_rows = CTAS == 2 ? N / 2 : N;                       // B rows held by this CTA
    const uint32_t idesc = make_idesc_mn(CTAS == 2 ? 256 : 128, N);
    if (tid == 0 && rank == 0) {
        // zero-init the accumulators with a first MMA, then time n_mma accumulating MMAs
        const long long t0 = clock64();
        const bool pat3 = variant & 16, commits = variant & 32;
        for (int i = 0; i < n_mma; ++i) {
            if (pat3) {   // the engines' pattern: per K-step a_hi*B_hi, a_lo*B_hi, a_hi*B_lo (A from TMEM, 8 KB per K-step per CTA-half)
                const int t = (i / 3) & 15, piece = i % 3;
                const uint32_t a = tmem + 16u * (uint32_t)t + (piece == 1 ? 8u : 0u);
                const uint32_t sb = sbase + 65536u + (uint32_t)(t & 3) * 8192u + (piece == 2 ? 4096u : 0u);
                const uint64_t bd = make_desc(sb, 2048, 128);
                if (CTAS == 2) mma_ts_pair(tmem + 256u, a, bd, idesc, i > 0); else mma_ts(tmem + 256u, a, bd, idesc, i > 0);
                if (commits && (i % 12) == 11) { if (CTAS == 2) tc_commit_pair(bar + 8, 1); else tc_commit(bar + 8); }
                continue;
            }
            const int t = same ? 0 : (i & 15);
            const uint32_t d = tmem + ((two_acc && (i & 1)) ? (a_smem ? 256u : 0u) : 0u) + (a_smem ? 0u : 256u);
            const uint64_t bd = make_desc(sbase + 65536u + (uint32_t)t * (uint32_t)(nb_rows * 32), (uint32_t)nb_rows * 16u, 128);
            if (a_smem) {
                const uint64_t ad = make_desc(sbase + (uint32_t)t * 4096u, 2048, 128);
                if (CTAS == 2) mma_ss_pair(d, ad, bd, idesc, i > 1); else mma_ss(d, ad, bd, idesc, i > 1);
            } else {
                const uint32_t a = tmem + 16u * (uint32_t)t;
                if (CTAS == 2) mma_ts_pair(d, a, bd, idesc, i > 0); else mma_ts(d, a, bd, idesc, i > 0);
            }
        }
        const long long t1 = clock64();
        if (CTAS == 2) tc_commit_pair(bar, 1); else tc_commit(bar);
        mbar_wait(bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = (unsigned long long)(t2 - t0); }
    }
    tc_fence_before();
    __syncthreads();
    if (CTAS == 2) cluster_sync_all();
    if (warp == 0) {
        if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int CTAS>
void run(int variant, int n_mma, int grid) {
    unsigned long long* d;
    cudaMalloc(&d, 16);
    auto k = rate_kernel<CTAS>;
    const int smem = 202 * 1024 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, k, variant, n_mma, d);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return; }
    }
    unsigned long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("cta_group %d  A %-4s  N %3d  %s  %s  grid %3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (%d MMAs)\n", CTAS,
           (variant & 1) ? "smem" : "tmem", (variant & 2) ? 128 : 256, (variant & 4) ? "2 acc" : "1 acc",
           (variant & 8) ? "same operands " : "16 K-step ring", grid, (double)h[0] / n_mma, (double)h[1] / n_mma, n_mma);
    cudaFree(d);
}

int main() {
    const int n = 960;
    for (int grid : {148}) {
        for (int v : {0, 2, 16, 18, 48, 50}) {
            run<1>(v, n, grid);
            run<2>(v, n, grid);
        }
    }
    return 0;
}
