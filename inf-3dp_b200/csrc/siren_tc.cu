// tcgen05 tensor-core engine of the fused SIREN jet (sm_100a): value or value+gradient of the flagship network
// 3 -> 256 -> 256 -> 256 -> 256 -> out (sine, out <= 8) in ONE persistent kernel, activations never leave the SM.
//
// Replaces modules.py:143-160 (forward) + diff_operators.py:128-132 (gradient via autograd) of the reference.
//
// Formulation (forward-mode jet).  A CTA owns 128 "rows" at a time:
//   ORDER 1: 32 points x 4 jet columns, row r = 4*p + c  (c = 0 value, 1..3 = d/dx, d/dy, d/dz)
//   ORDER 0: 128 points, row r = p
// A hidden layer is one 128 x 256 x 256 GEMM  D[r, n] = sum_k A[r, k] * W'[n, k]  on the 5th-gen tensor cores:
//   * A (activations) lives in TENSOR MEMORY and is consumed from there (tcgen05.mma, A-from-TMEM form),
//     B (weights, FP16 UMMA K-major core-matrix images prepared at pack time) streams L2 -> shared memory through
//     a 6-stage cp.async.bulk / mbarrier ring, D accumulates in FP32 in the other half of tensor memory.
//   * PRECISE mode splits both operands into FP16 hi + lo and issues three MMAs per K-step
//     (a_hi*W_hi + a_lo*W_hi + a_hi*W_lo): FP32-class accuracy (measured need, SURVEY.md section 7 hard part 1);
//     FAST mode issues a_hi*W_hi only.
//   * The epilogue warps read D with tcgen05.ld, apply the jet update
//         value row:   a  = sin(z + b)            tangent row: da = cos(z_value + b) * dz
//     (the tangent lanes fetch their point's z with one quad shuffle and evaluate cos as sin(. + pi/2), so all 32
//     lanes run one MUFU.SIN per feature), split to FP16 hi/lo and write the next layer's A operand with tcgen05.st
//     IN PLACE over the accumulator columns they just consumed.  The two halves of tensor memory swap roles every
//     layer; the loop over K is the outer MMA loop (N = 256 per instruction), so the MMAs of layer l+1 start as
//     soon as the first 16-feature slice of layer l's activations is written and overlap the rest of the epilogue.
//   * Layer 0 (3 -> 256) and the omega0 folding are FP32 CUDA-core work (north_star), the output layer is a
//     128 x 16 x 256 MMA.
// Warp roles (576 threads): warps 0-15 epilogue (four warps per TMEM lane quadrant, interleaved 16-feature
// slices = one MMA K-step each), warp 16 weight producer (one elected lane), warp 17 TMEM allocator + MMA issuer
// (one elected lane).  Within a 256-column half of tensor memory, K-step t of the A operand lives in columns
// [16t, 16t+8) (FP16 hi, two features per column) and [16t+8, 16t+16) (FP16 lo): exactly the 16 FP32 accumulator
// columns the slice was computed from, which is what makes the in-place conversion hazard-free.
// Every mbarrier wait carries a watchdog that traps instead of hanging the GPU.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace inf3dp {

namespace {

// ---- packed TC section layout (offsets relative to SirenHeader::off_tc) -----------------------------------
constexpr size_t kTcOffScales = 0;        // float inv_scale[4]: hidden layers 1..3, output layer
constexpr size_t kTcOffB4 = 64;           // float b4[8]
constexpr size_t kTcOffMaxAbs = 128;      // uint32 max |w| bits per layer [4] (pack-time scratch)
constexpr size_t kTcOffBias = 256;        // float bias_phase[3][256][2] = (b', b' + pi/2)
constexpr size_t kTcOffOutImg = 7168;     // out-layer image: 16 K-steps x 512 B hi, then lo  (16 KB)
constexpr size_t kTcOffHidImg = 7168 + 16384;  // 3 layers x 8 stages x 32 KB
constexpr size_t kStageBytes = 32768;     // 2 K-steps x (hi 8 KB) then 2 K-steps x (lo 8 KB)
constexpr size_t kLayerImgBytes = 8 * kStageBytes;
constexpr size_t kTcSectionBytes = kTcOffHidImg + 3 * kLayerImgBytes;

// ---- shared memory layout -----------------------------------------------------------------------------------
constexpr int kStages = 3;                 // ring stages of 4 K-steps (2 consecutive 32 KB image stages)
constexpr uint32_t kRingStageBytes = 2 * kStageBytes;
constexpr uint32_t kSmemRing = 0;
constexpr uint32_t kSmemW4 = kStages * kRingStageBytes;              // 196608: f32 W4[4][256] (output layer)
constexpr uint32_t kSmemPartial = kSmemW4 + 4096;                    // f32 partial[4 warp groups][4 outputs][128 rows]
constexpr uint32_t kSmemBias = kSmemPartial + 8192;
constexpr uint32_t kSmemW0 = kSmemBias + 6144;                       // 219136
constexpr uint32_t kSmemB0 = kSmemW0 + 4096;                         // 223232
constexpr uint32_t kSmemMisc = kSmemB0 + 2048;                       // inv_scale[4], b4[8]   (b0 table is [256][2])
constexpr uint32_t kSmemBars = kSmemMisc + 64;                       // 224320
constexpr int kBarWFull = 0, kBarWEmpty = kStages, kBarAReady = 2 * kStages, kBarDFull = 2 * kStages + 4,
              kNumBars = 2 * kStages + 5;
constexpr uint32_t kSmemTmemPtr = kSmemBars + kNumBars * 8;
constexpr uint32_t kSmemTotal = kSmemTmemPtr + 16;
static_assert(kSmemTotal <= 232448 - 1024, "shared memory budget (1 KB reserved for base alignment)");

constexpr int kEpiWarps = 16;
constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
constexpr int kThreadsTc = (kEpiWarps + 2) * 32;  // 576

using namespace tcptx;

// ---------------------------------------------------------------------------------------------------------------
template <int ORDER, bool PRECISE, int OUTC>
__global__ void __launch_bounds__(kThreadsTc, 1)
siren_tc_kernel(const char* __restrict__ packed, SirenHeader h, const float* __restrict__ x, int64_t n,
                float* __restrict__ y, float* __restrict__ J, unsigned long long* __restrict__ dbg) {
    extern __shared__ unsigned char smem_dyn[];
    // 1 KB alignment for the bulk-copy destinations / UMMA operands
    const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* sptr = smem_dyn + (sbase - smem_u32(smem_dyn));

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const char* tc = packed + h.off_tc;

    constexpr int kPts = ORDER == 1 ? 32 : 128;   // points per tile
    const int out_f = h.out_features;
    const int64_t num_tiles = (n + kPts - 1) / kPts;
    const int my_tiles = (int)((num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);  // blockIdx.x < num_tiles

    const uint32_t bars = sbase + kSmemBars;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

    // ---- one-time setup --------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar(kBarWFull + s), 1); mbar_init(bar(kBarWEmpty + s), 1); }
        for (int r = 0; r < 4; ++r) mbar_init(bar(kBarAReady + r), kEpiWarps);  // round r: every warp's r-th slice
        mbar_init(bar(kBarDFull), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // small FP32 tables through the generic proxy (read by CUDA cores only)
        float* s_bias = reinterpret_cast<float*>(sptr + kSmemBias);
        const float* g_bias = reinterpret_cast<const float*>(tc + kTcOffBias);
        for (int i = tid; i < 3 * 256 * 2; i += kThreadsTc) s_bias[i] = g_bias[i];
        float* s_w0 = reinterpret_cast<float*>(sptr + kSmemW0);
        const float* g_w0 = reinterpret_cast<const float*>(packed + h.off_w0);
        // [f] = (w_x, w_y, w_z, 1): lane c reads its layer-0 tangent factor as element (c + 3) & 3 -- no branches
        for (int i = tid; i < 256 * 4; i += kThreadsTc) s_w0[i] = (i & 3) == 3 ? 1.f : g_w0[i];
        float* s_b0 = reinterpret_cast<float*>(sptr + kSmemB0);
        const float* g_b0 = reinterpret_cast<const float*>(packed + h.off_b0);
        for (int i = tid; i < 256; i += kThreadsTc) { s_b0[2 * i] = g_b0[i]; s_b0[2 * i + 1] = g_b0[i] + kHalfPi; }
        float* s_w4 = reinterpret_cast<float*>(sptr + kSmemW4);
        const float* g_w4 = reinterpret_cast<const float*>(packed + h.off_wout);  // [kMaxOut][256], rows >= out are 0
        for (int i = tid; i < 4 * 256; i += kThreadsTc) s_w4[i] = g_w4[i];
        float* s_misc = reinterpret_cast<float*>(sptr + kSmemMisc);
        if (tid < 4) s_misc[tid] = reinterpret_cast<const float*>(tc + kTcOffScales)[tid];
        if (tid >= 4 && tid < 12) s_misc[tid] = reinterpret_cast<const float*>(tc + kTcOffB4)[tid - 4];
    }
    if (warp == kMmaWarp) {  // TMEM: all 512 columns (one CTA per SM by construction: ~219 KB of shared memory)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(sbase + kSmemTmemPtr), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sptr + kSmemTmemPtr);
    const uint32_t regionP = tmem, regionQ = tmem + 256;

    if (warp == kProducerWarp) {
        // ================= weight producer =====================================================================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < my_tiles; ++it) {
                for (int l = 0; l < 3; ++l) {
                    const char* img = tc + kTcOffHidImg + (size_t)l * kLayerImgBytes;
                    for (int s = 0; s < 4; ++s) {  // 4 K-steps per ring stage
                        mbar_wait(bar(kBarWEmpty + stage), phase ^ 1u);
                        const uint32_t dst = sbase + kSmemRing + stage * kRingStageBytes;
                        const char* src = img + (size_t)s * kRingStageBytes;
                        if (PRECISE) {
                            mbar_expect_tx(bar(kBarWFull + stage), kRingStageBytes);
                            bulk_g2s(dst, src, kRingStageBytes, bar(kBarWFull + stage));
                        } else {  // FP16 hi halves only (first 16 KB of each 32 KB image stage)
                            mbar_expect_tx(bar(kBarWFull + stage), 32768u);
                            bulk_g2s(dst, src, 16384u, bar(kBarWFull + stage));
                            bulk_g2s(dst + 32768u, src + 32768, 16384u, bar(kBarWFull + stage));
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer: the whole warp runs the loop convergently, one elected lane issues (tc_ptx.cuh) ==========
        {
            constexpr uint32_t idesc256 = make_idesc(256);
            uint32_t stage = 0, wphase = 0;
            uint32_t aphase = 0;  // parity of the a_ready barriers (one completion per activation version)
            unsigned long long wait_a = 0, wait_w = 0, dbg_lag = 0;
            uint32_t dbg_dphase = 0;
            unsigned long long* pa = dbg ? &wait_a : nullptr;
            unsigned long long* pw = dbg ? &wait_w : nullptr;
            const long long t_begin = clock64();
            for (int it = 0; it < my_tiles; ++it) {
                // hidden layers 1..3: A alternates P,Q,P ; D alternates Q,P,Q
                for (int l = 0; l < 3; ++l) {
                    const uint32_t regA = (l & 1) ? regionQ : regionP;
                    const uint32_t regD = (l & 1) ? regionP : regionQ;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        mbar_wait_timed(bar(kBarAReady + r), aphase, pa);   // K-steps 4r..4r+3 of the activations
                        mbar_wait_timed(bar(kBarWFull + stage), wphase, pw);
                        tc_fence_after();
                        const uint32_t ring = sbase + kSmemRing + stage * kRingStageBytes;
#pragma unroll
                        for (int tt = 0; tt < 4; ++tt) {
                            const int t = 4 * r + tt;
                            const uint32_t sb = ring + (uint32_t)(tt >> 1) * 32768u + (uint32_t)(tt & 1) * 8192u;
                            const uint64_t bh = make_desc(sb, 4096, 128);
                            const uint32_t a_hi = regA + 16u * (uint32_t)t;
                            mma_ts_elect(regD, a_hi, bh, idesc256, t > 0 ? 1u : 0u);
                            if (PRECISE) {
                                const uint64_t bl = make_desc(sb + 16384u, 4096, 128);
                                mma_ts_elect(regD, a_hi + 8u, bh, idesc256, 1u);
                                mma_ts_elect(regD, a_hi, bl, idesc256, 1u);
                            }
                        }
                        tc_commit_elect(bar(kBarWEmpty + stage));  // stage free once these MMAs have read it
                        if (++stage == kStages) { stage = 0; wphase ^= 1u; }
                    }
                    tc_commit_elect(bar(kBarDFull));
                    if (dbg) {  // instrumentation only: how long after the last issue does D become visible?
                        const long long t0 = clock64();
                        mbar_wait(bar(kBarDFull), dbg_dphase);
                        dbg_lag += (unsigned long long)(clock64() - t0);
                    }
                    dbg_dphase ^= 1u;
                    aphase ^= 1u;
                }
                // (the linear output layer is folded into the last hidden layer's epilogue on the CUDA cores)
            }
            if (dbg && blockIdx.x == 0 && lane == 0) {
                dbg[4] = dbg_lag;
                dbg[0] = (unsigned long long)(clock64() - t_begin);
                dbg[1] = wait_a;
                dbg[2] = wait_w;
                dbg[3] = (unsigned long long)my_tiles;
            }
        }
    } else {
        // ================= epilogue warps 0..15 ==================================================================
        const int quad = warp & 3;           // TMEM lane quadrant this warp may access
        const int grp = warp >> 2;           // slices grp, grp+4, grp+8, grp+12 (one MMA K-step each)
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const int row = quad * 32 + lane;    // TMEM lane = tile row
        const int c = ORDER == 1 ? (lane & 3) : 0;          // jet column of this row
        const int p_local = ORDER == 1 ? (row >> 2) : row;  // point within the tile
        const int qbase = lane & ~3;                         // lane holding this point's value row
        const float* s_bias = reinterpret_cast<const float*>(sptr + kSmemBias) + (c != 0 ? 1 : 0);
        const float4* s_w0 = reinterpret_cast<const float4*>(sptr + kSmemW0);
        const float* s_w0m = reinterpret_cast<const float*>(sptr + kSmemW0) + ((c + 3) & 3);  // 1, w_x, w_y, w_z
        const float* s_b0 = reinterpret_cast<const float*>(sptr + kSmemB0) + (c != 0 ? 1 : 0); // b0 (+ pi/2)
        const float* s_misc = reinterpret_cast<const float*>(sptr + kSmemMisc);
        uint32_t dphase = 0;
        unsigned long long wait_d = 0;
        unsigned long long* pd = (dbg && lane == 0) ? &wait_d : nullptr;

        for (int it = 0; it < my_tiles; ++it) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const int64_t pt = tile * kPts + p_local;
            const bool pvalid = pt < n;
            float px = 0.f, py = 0.f, pz = 0.f;
            if (pvalid) { px = x[3 * pt]; py = x[3 * pt + 1]; pz = x[3 * pt + 2]; }

            // ---- layer 0 (FP32): a0 = sin(W0' x + b0'), da0_i = cos(.) * W0'[:, i]  -> region P ----------------
#pragma unroll 1
            for (int mi = 0; mi < 4; ++mi) {
                const int m = grp + 4 * mi;
                uint32_t o[16];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    float r2[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int f = 16 * m + i + u;
                        const float4 w = s_w0[f];
                        float z = fmaf(w.x, px, s_b0[2 * f]);   // b0' (+ pi/2 on tangent rows: cos = sin(. + pi/2))
                        z = fmaf(w.y, py, z);
                        z = fmaf(w.z, pz, z);
                        r2[u] = sin_layer0(z) * s_w0m[4 * f];
                    }
                    split_pack<PRECISE>(r2[0], r2[1], o[i >> 1], o[8 + (i >> 1)]);
                }
                const uint32_t addr = regionP + lane_off + 16u * (uint32_t)m;
                if (PRECISE) { TC_ST16(addr, o); } else { TC_ST8(addr, o); }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(kBarAReady + mi));
            }

            // ---- hidden layers 1..3: D (FP32, region regD) -> activations in place --------------------------------
            // per-lane constants so that value and tangent rows run the same instruction stream:
            //   out = sin(z_value * inv + b (+ pi/2)) * (own * k1 + k0),  (k1,k0) = (0,1) value row, (inv,0) tangent row
            // The LAST hidden layer does not write activations back: the linear output layer
            //   y = W4 a3 + b4,   dy/dx_i = W4 da3_i        (exact FP32 weights)
            // is accumulated on the CUDA cores while the activations are in registers.
            float acc[OUTC];
#pragma unroll
            for (int o = 0; o < OUTC; ++o) acc[o] = 0.f;
#pragma unroll 1
            for (int l = 0; l < 3; ++l) {
                const uint32_t regD = (l & 1) ? regionP : regionQ;
                const float inv_scale = s_misc[l];
                const float k1 = (ORDER == 1 && c != 0) ? inv_scale : 0.f;
                const float k0 = (ORDER == 1 && c != 0) ? 0.f : 1.f;
                const float* bl = s_bias + l * 512;
                mbar_wait_timed(bar(kBarDFull), dphase, pd);   // every lane polls (one polling lane + __syncwarp wakes the warp later)
                dphase ^= 1u;
                tc_fence_after();
                uint32_t v[16];
                TC_LD16(regD + lane_off + 16u * (uint32_t)grp, v);
#pragma unroll 1
                for (int mi = 0; mi < 4; ++mi) {
                    const int m = grp + 4 * mi;
                    const uint32_t addr = regD + lane_off + 16u * (uint32_t)m;
                    tc_wait_ld();
                    float cur[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) cur[i] = __uint_as_float(v[i]);
                    if (mi < 3) TC_LD16(addr + 64u, v);  // prefetch the next slice (columns 16*(m+4)) under the math
                    if (l < 2) {
                        uint32_t o[16];
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            float r2[2];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int f = 16 * m + i + u;
                                const float own = cur[i + u];
                                float zv = own;
                                if (ORDER == 1) zv = __shfl_sync(0xffffffffu, own, qbase);
                                const float arg = fmaf(zv, inv_scale, bl[2 * f]);  // + b' (+ pi/2 on tangent rows)
                                const float sn = sin_mufu(arg);
                                r2[u] = ORDER == 1 ? sn * fmaf(own, k1, k0) : sn;
                            }
                            split_pack<PRECISE>(r2[0], r2[1], o[i >> 1], o[8 + (i >> 1)]);
                        }
                        if (PRECISE) { TC_ST16(addr, o); } else { TC_ST8(addr, o); }
                        tc_wait_st();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar(kBarAReady + mi));
                    } else {
                        const float* s_w4 = reinterpret_cast<const float*>(sptr + kSmemW4);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int f = 16 * m + i;
                            const float own = cur[i];
                            float zv = own;
                            if (ORDER == 1) zv = __shfl_sync(0xffffffffu, own, qbase);
                            const float arg = fmaf(zv, inv_scale, bl[2 * f]);
                            const float sn = sin_mufu(arg);
                            const float val = ORDER == 1 ? sn * fmaf(own, k1, k0) : sn;
#pragma unroll
                            for (int o = 0; o < OUTC; ++o) acc[o] = fmaf(val, s_w4[o * 256 + f], acc[o]);
                        }
                    }
                }
            }

            // ---- output: sum the four feature-quarter partials of every row in a fixed order ------------------------
            {
                float* s_part = reinterpret_cast<float*>(sptr + kSmemPartial);
#pragma unroll
                for (int o = 0; o < OUTC; ++o) s_part[(grp * 4 + o) * 128 + row] = acc[o];
                asm volatile("bar.sync 1, 512;" ::: "memory");  // the 16 epilogue warps only
                if (grp == 0 && pvalid) {
#pragma unroll
                    for (int o = 0; o < OUTC; ++o) {
                        if (o < out_f) {
                            const float val = ((s_part[(0 * 4 + o) * 128 + row] + s_part[(1 * 4 + o) * 128 + row]) +
                                               s_part[(2 * 4 + o) * 128 + row]) + s_part[(3 * 4 + o) * 128 + row];
                            if (c == 0) y[pt * out_f + o] = val + s_misc[4 + o];
                            else J[(pt * out_f + o) * 3 + (c - 1)] = val;
                        }
                    }
                }
            }
            // Hazards into the next tile: region P (next layer-0 target) was last READ by the layer-3 MMAs, which
            // completed before this warp passed d_full(3); region Q and s_part are only touched again behind
            // a_ready / d_full barriers that need every epilogue warp to have left this point.
        }
        if (dbg && blockIdx.x == 0 && lane == 0) dbg[8 + warp] = wait_d;
    }

    // ---- teardown ------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// ---- pack-time kernels ---------------------------------------------------------------------------------------------
__global__ void tc_maxabs_kernel(const float* __restrict__ W, int count, uint32_t* __restrict__ slot) {
    float m = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(W[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(m));
}

// power-of-two scale that puts max |w * mul| into [8, 16): FP16 hi AND lo parts stay normal numbers
__device__ __forceinline__ float pow2_scale(float maxabs) {
    if (!(maxabs > 0.f) || !isfinite(maxabs)) return 1.f;
    int e;
    frexpf(maxabs, &e);            // maxabs = m * 2^e, m in [0.5, 1)
    return ldexpf(1.f, 4 - e);     // maxabs * scale in [8, 16)
}

__global__ void tc_pack_hidden_kernel(const float* __restrict__ W, const float* __restrict__ b, float omega0, int layer,
                                      char* __restrict__ tc) {
    const uint32_t* mx = reinterpret_cast<const uint32_t*>(tc + kTcOffMaxAbs);
    const float scale = pow2_scale(omega0 * __uint_as_float(mx[layer]));
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) reinterpret_cast<float*>(tc + kTcOffScales)[layer] = 1.f / scale;
    if (idx < 256) {
        float* bp = reinterpret_cast<float*>(tc + kTcOffBias) + (size_t)layer * 512 + 2 * idx;
        const float bb = omega0 * b[idx];
        bp[0] = bb;
        bp[1] = bb + kHalfPi;
    }
    if (idx >= 256 * 256) return;
    const int nrow = idx >> 8, k = idx & 255;
    const float v = omega0 * W[idx] * scale;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const int stage = k >> 5, j = (k >> 4) & 1, kc = (k >> 3) & 1, ng = nrow >> 3;
    const size_t off = (size_t)layer * kLayerImgBytes + (size_t)stage * kStageBytes + (size_t)j * 8192 + (size_t)kc * 4096 +
                       (size_t)ng * 128 + (size_t)(nrow & 7) * 16 + (size_t)(k & 7) * 2;
    char* img = tc + kTcOffHidImg;
    *reinterpret_cast<__half*>(img + off) = hi;
    *reinterpret_cast<__half*>(img + off + 16384) = lo;
}

__global__ void tc_pack_out_kernel(const float* __restrict__ W, const float* __restrict__ b, int out_f,
                                   char* __restrict__ tc) {
    const uint32_t* mx = reinterpret_cast<const uint32_t*>(tc + kTcOffMaxAbs);
    const float scale = pow2_scale(__uint_as_float(mx[3]));
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) reinterpret_cast<float*>(tc + kTcOffScales)[3] = 1.f / scale;
    if (idx < 8) reinterpret_cast<float*>(tc + kTcOffB4)[idx] = idx < out_f ? b[idx] : 0.f;
    if (idx >= 16 * 256) return;
    const int nrow = idx >> 8, k = idx & 255;
    const float v = nrow < out_f ? W[nrow * 256 + k] * scale : 0.f;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const int t = k >> 4, kc = (k >> 3) & 1, ng = nrow >> 3;
    const size_t off = (size_t)t * 512 + (size_t)kc * 256 + (size_t)ng * 128 + (size_t)(nrow & 7) * 16 + (size_t)(k & 7) * 2;
    char* img = tc + kTcOffOutImg;
    *reinterpret_cast<__half*>(img + off) = hi;
    *reinterpret_cast<__half*>(img + off + 8192) = lo;
}

template <int ORDER, bool PRECISE, int OUTC>
int launch_tc(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, float* J,
              cudaStream_t stream) {
    auto kern = siren_tc_kernel<ORDER, PRECISE, OUTC>;
    const int smem = (int)kSmemTotal + 1024;
    INF3DP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    constexpr int kPts = ORDER == 1 ? 32 : 128;
    const int64_t tiles = (n + kPts - 1) / kPts;
    const int grid = (int)(tiles < (int64_t)siren_sms() ? tiles : (int64_t)siren_sms());
    static const bool debug = getenv("INF3DP_TC_DEBUG") != nullptr;
    unsigned long long* dbg = nullptr;
    if (debug) {
        INF3DP_CUDA(cudaMalloc(&dbg, 64 * sizeof(unsigned long long)));
        INF3DP_CUDA(cudaMemsetAsync(dbg, 0, 64 * sizeof(unsigned long long), stream));
    }
    kern<<<grid, kThreadsTc, smem, stream>>>((const char*)packed, h, x, n, y, J, dbg);
    INF3DP_LAUNCH_CHECK("siren_tc_kernel");
    if (debug) {
        unsigned long long hbuf[64];
        INF3DP_CUDA(cudaMemcpyAsync(hbuf, dbg, sizeof(hbuf), cudaMemcpyDeviceToHost, stream));
        INF3DP_CUDA(cudaStreamSynchronize(stream));
        INF3DP_CUDA(cudaFree(dbg));
        const double tiles = hbuf[3] ? (double)hbuf[3] : 1.0;
        fprintf(stderr, "[inf3dp tc debug] block0: %.0f tiles, %.0f cyc/tile; MMA warp blocked on a_ready %.0f, on w_full %.0f cyc/tile; "
                        "epilogue warp0 blocked on d_full %.0f, warp5 %.0f, warp15 %.0f cyc/tile; issue->D visible lag %.0f cyc/tile (3 hidden layers)\n",
                tiles, hbuf[0] / tiles, hbuf[1] / tiles, hbuf[2] / tiles, hbuf[8] / tiles, hbuf[13] / tiles, hbuf[23] / tiles, hbuf[4] / tiles);
    }
    return INF3DP_OK;
}

}  // namespace

size_t siren_tc_section_bytes() { return kTcSectionBytes; }

const uint32_t* siren_tc_maxabs_bits(const void* packed, const SirenLayout& lay) {
    return reinterpret_cast<const uint32_t*>((const char*)packed + lay.off_tc + kTcOffMaxAbs);
}

int siren_tc_pack(const float* const* W, const float* const* b, float omega0, int out_features, void* packed,
                  const SirenLayout& lay, SirenHeader& hdr, cudaStream_t stream) {
    char* tc = (char*)packed + lay.off_tc;
    INF3DP_CUDA(cudaMemsetAsync(tc, 0, kTcOffHidImg, stream));  // scales, max-abs scratch, out image padding rows
    uint32_t* mx = reinterpret_cast<uint32_t*>(tc + kTcOffMaxAbs);
    for (int l = 0; l < 3; ++l) {
        tc_maxabs_kernel<<<64, 256, 0, stream>>>(W[1 + l], 256 * 256, mx + l);
        INF3DP_LAUNCH_CHECK("tc_maxabs_kernel");
    }
    tc_maxabs_kernel<<<4, 256, 0, stream>>>(W[4], out_features * 256, mx + 3);
    INF3DP_LAUNCH_CHECK("tc_maxabs_kernel");
    for (int l = 0; l < 3; ++l) {
        tc_pack_hidden_kernel<<<256, 256, 0, stream>>>(W[1 + l], b[1 + l], omega0, l, tc);
        INF3DP_LAUNCH_CHECK("tc_pack_hidden_kernel");
    }
    tc_pack_out_kernel<<<16, 256, 0, stream>>>(W[4], b[4], out_features, tc);
    INF3DP_LAUNCH_CHECK("tc_pack_out_kernel");
    hdr.tc_ok = 1;
    return INF3DP_OK;
}

int siren_tc_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, int order, float* y,
                    float* J, bool precise, cudaStream_t stream) {
    INF3DP_CHECK_ARG(h.tc_ok && h.off_tc != 0, "siren_tc: packed weights carry no tensor-core section");
    const bool one = h.out_features == 1;
    if (order == 0) {
        if (precise) return one ? launch_tc<0, true, 1>(packed, h, x, n, y, J, stream) : launch_tc<0, true, 4>(packed, h, x, n, y, J, stream);
        return one ? launch_tc<0, false, 1>(packed, h, x, n, y, J, stream) : launch_tc<0, false, 4>(packed, h, x, n, y, J, stream);
    }
    if (precise) return one ? launch_tc<1, true, 1>(packed, h, x, n, y, J, stream) : launch_tc<1, true, 4>(packed, h, x, n, y, J, stream);
    return one ? launch_tc<1, false, 1>(packed, h, x, n, y, J, stream) : launch_tc<1, false, 4>(packed, h, x, n, y, J, stream);
}

}  // namespace inf3dp
