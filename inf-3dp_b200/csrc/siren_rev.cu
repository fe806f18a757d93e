// tcgen05 REVERSE-mode engine of the fused SIREN value + gradient query (sm_100a), single-output networks
// 3 -> 256 -> 256 -> 256 -> 256 -> 1 (sine): the SDF / slice / guidance fields of INF-3DP.
//
// Replaces modules.py:143-160 (forward) + diff_operators.py:128-132 (gradient via autograd) of the reference -- and it
// evaluates exactly what the reference's autograd evaluates (one forward sweep, one backward sweep), but inside ONE
// persistent kernel with every activation on chip:
//
//   forward   z_l = a_{l-1} W_l'^T + b_l'     a_l = sin z_l     c_l = cos z_l          l = 1..3   (W' = 30 W, b' = 30 b)
//   output    y   = a_3 . w4 + b4             g_3 = w4 (.) c_3                                     (dy/dz_3)
//   backward  g_{l-1} = (g_l W_l') (.) c_{l-1}                                            l = 3..1
//   input     dy/dx_i = sum_f g_0[f] W_0'[f, i]
//
// A tile is 128 POINTS (row = point = one lane of tensor memory), so value + gradient cost 6 GEMMs of 128x256x256 per
// 128 points -- half of the forward-mode jet (3 GEMMs per 32 points) -- and one sin + one cos per point-feature
// instead of four MUFU evaluations.  Algorithmic work: 2 * (2*3*256 + 2*3*256^2 + 2*256) = 790 528 FLOP / point
// (the "reverse-mode equivalent" figure).
//
// Execution (CTAS = 2, the shipped configuration): a CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns 256 points.
//   * Each GEMM is one M=256 x N=256 x K=16 UMMA per K-step and operand piece, issued by one thread of the leader CTA:
//     A (activations or upstream gradients, FP16 hi + lo) is read from each CTA's TENSOR MEMORY, B (weights) from
//     shared memory where each CTA holds only ITS 128 of the 256 B rows -- the L2 -> SM weight stream per SM is half
//     of the single-CTA form, which is what the single-CTA forward engine is limited by -- and D accumulates in FP32 in
//     the other 256 columns of tensor memory.  Three MMAs per K-step (a_hi W_hi + a_lo W_hi + a_hi W_lo): the measured
//     accuracy need (scripts/emu_reverse.py: any unsplit operand, forward or backward, breaks the 0.1 degree bound).
//   * 16 epilogue warps per CTA turn D into the next A operand IN PLACE (tcgen05.ld, FP32 math, FP16 hi/lo split,
//     tcgen05.st over the same columns); the two halves of tensor memory swap roles after every GEMM.
//   * cos z_1, cos z_2 are needed again by the backward sweep: they are stashed in FP32 (an FP16 stash fails parity) in a
//     per-CTA scratch that stays in L2 (coalesced 512 B per warp store); cos z_3 is consumed immediately, cos z_0 is
//     recomputed from the point in FP32.
//   * Weight images stream L2 -> shared memory through a cp.async.bulk / mbarrier ring; the peer CTA forwards its
//     "stage landed" to the leader with a remote mbarrier arrive; tcgen05.commit multicasts "stage free" and
//     "accumulator ready" to both CTAs.
// CTAS = 1 (developer configuration, INF3DP_REV_CTAS=1) runs the same schedule on one CTA with two N=128 MMAs.
// Every mbarrier wait carries a watchdog that traps instead of hanging the GPU.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace inf3dp {

namespace {

using namespace tcptx;

// ---- packed REV section (offsets relative to SirenHeader::off_rev) ------------------------------------------------
constexpr size_t kRevOffMisc = 0;           // float: G, 1/G
constexpr size_t kRevOffW4 = 1024;          // float2[256]: (w4[f], G * w4[f])
constexpr size_t kRevOffImg = 4096;         // 6 GEMMs x 2 N-halves x 16 K-steps x (hi 4 KB, lo 4 KB)
constexpr size_t kPerKBytes = 8192;         // one K-step of one N-half: hi then lo, each 2 K-halves x 128 rows x 16 B
constexpr size_t kHalfImgBytes = 16 * kPerKBytes;
constexpr size_t kGemmImgBytes = 2 * kHalfImgBytes;
constexpr size_t kRevSectionBytes = kRevOffImg + 6 * kGemmImgBytes;

// packed MLP section (level classifier 4 -> 256 -> 256 -> C, ReLU; same off_rev slot of the blob)
constexpr size_t kMlpOffMisc = 0;           // float: [0] 1/S, uint32 at +64: max |W1| bits (pack-time scratch), float b2[16] at +128
constexpr size_t kMlpOffW2 = 1024;          // float [256][16]: W2 transposed, columns >= C are zero
constexpr size_t kMlpOffImg = 1024 + 16384; // one GEMM image (W1), 2 N-halves x 16 K-steps x (hi 4 KB, lo 4 KB)
constexpr size_t kMlpSectionBytes = kMlpOffImg + 2 * 16 * 8192;
constexpr int kMlpMaxOut = 16;

constexpr int kEpiWarps = 16;
constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1, kFirstEpiWarp = 0;
constexpr int kThreads = (kEpiWarps + 2) * 32;  // 576
constexpr int kStageK = 4;                       // K-steps per ring stage
constexpr size_t kStashBytesPerCta = 2 * 16 * 4 * 128 * 16;  // [2 layers][16 slices][4][128 rows] float4 = 256 KB

template <int CTAS, int KIND = 0>
struct RevCfg {
    // weight ring: 32 KB stages of four K-steps (a six-stage ring was measured: 494 instead of 504 Mpts/s on the 512^3 grid, the
    // MMA thread's wait for weights unchanged -- it is the peer CTA's relay, not the depth)
    static constexpr int kStages = CTAS == 2 ? 4 : 3;
    static constexpr uint32_t kStageBytes = (uint32_t)(kStageK * kPerKBytes) * (CTAS == 2 ? 1u : 2u);
    static constexpr uint32_t kSmemRing = 0;
    static constexpr uint32_t kSmemW0 = kStages * kStageBytes;   // float4[256] (wx, wy, wz, b0)
    static constexpr uint32_t kSmemBias = kSmemW0 + 4096;        // float[3][256]
    static constexpr uint32_t kSmemW4 = kSmemBias + 3072;        // float2[256]
    static constexpr uint32_t kSmemPart = kSmemW4 + 2048;        // float[4 groups][4][128]
    static constexpr uint32_t kSmemNext = kSmemPart + 8192;      // float[512][3]: the next tile's point, one slot per epilogue thread
    static constexpr uint32_t kSmemMisc = kSmemNext + 6144;      // inv_scale[3], b4, G, 1/G
    static constexpr uint32_t kSmemBars = kSmemMisc + 64;
    // a_ready: one barrier per operand round (KIND 1, 2: four rounds of four K-steps) or per K-step (KIND 0, 3: sixteen)
    static constexpr int kBarWFull = 0, kBarWEmpty = kStages, kBarAReady = 2 * kStages, kBarDFull = 2 * kStages + 16,
                         kNumBars = 2 * kStages + 17;
    static constexpr uint32_t kSmemTmemPtr = kSmemBars + kNumBars * 8;
    static constexpr uint32_t kSmemTotal = kSmemTmemPtr + 16;
    // KIND 2 only: W2^T [256][16] f32 and the partial logits [4 groups][16][128] f32
    static constexpr uint32_t kSmemMlpW2 = (kSmemTotal + 255u) & ~255u;
    static constexpr uint32_t kSmemMlpPart = kSmemMlpW2 + 16384;
    static constexpr uint32_t kSmemTotalMlp = kSmemMlpPart + 32768;
    static_assert(CTAS == 1 || (KIND == 2 ? kSmemTotalMlp : kSmemTotal) <= 232448 - 1024, "shared memory budget");   // KIND 2 runs on CTA pairs only
};

// KIND 0: reverse-mode value + gradient (6 GEMMs per tile of 128 points).
// KIND 1: forward-mode second-order jet, value + gradient + Hessian (diff_operators.py:27-44 hessian3d): 3 GEMMs per tile
//         of 12 points x 10 jet rows (value, d/dx_i, d2/dx_i dx_j for ij in 00 01 02 11 12 22; 3 points per 32-lane
//         quadrant of tensor memory, lanes 30 and 31 idle), same CTA-pair schedule, same weight images (GEMMs 0..2).
// KIND 2: the level classifier of the collision query (exp_scripts/train_levels.py:35-51, utils.py:269-303): ReLU MLP
//         4 -> 256 -> 256 -> C (C <= 16), value only.  One GEMM per tile of 128 points; layer 0 (4 -> 256) and the linear
//         output layer (256 -> C, folded into the hidden layer's epilogue) are FP32 CUDA-core work.
// KIND 3: value + gradient + Hessian by FORWARD-over-REVERSE differentiation (diff_operators.py:6-44 builds the same thing
//         with a second autograd pass): the KIND 0 sweep on a tile of 32 points x 4 rows -- the point itself and its three
//         tangents d/dx_i carried through the forward AND the backward sweep (6 GEMMs, same weight images).  Row 4p + c of
//         tensor memory: c = 0 the primal row (a_l, then G g_l), c = 1..3 the tangent rows (da_l/dx_i, then kT G dg_l/dx_i).
//         The primal row yields (y, dy/dx), tangent row i yields row i of the Hessian.  24 row-GEMMs per point instead of the
//         30 of the second-order jet, and about half of its epilogue instructions per point.
// KIND 4: value only (order 0) of the same networks -- the forward sweep of KIND 0 alone: 3 GEMMs per tile of 128 points, no
//         cosines, no scratch, the column-split epilogue; the unit boundary is fused as in KIND 1 (odd number of GEMMs per unit:
//         the two halves of tensor memory swap roles from unit to unit).
// Destinations of the packed (value, grad) rows when the query is fused with the multi-GPU all-gather (KIND 0 only):
// either one NVLink multicast address (multimem.st, replicated by the switch) or up to 8 peer-mapped buffers.
constexpr int kMaxRowPeers = 8;
struct RowPeers {
    float4* dst[kMaxRowPeers];
    float4* mc;
    int n;
};

template <int CTAS, int KIND>
__global__ void __launch_bounds__(kThreads, 1)
siren_rev_kernel(const char* __restrict__ packed, SirenHeader h, const float* __restrict__ x, int64_t n,
                 float* __restrict__ y, float* __restrict__ J, float* __restrict__ Hout, float4* __restrict__ stash,
                 unsigned long long* __restrict__ dbg, const RowPeers peers) {
    using C = RevCfg<CTAS, KIND>;
    constexpr int kGemms = (KIND == 0 || KIND == 3) ? 6 : ((KIND == 1 || KIND == 4) ? 3 : 1);
    constexpr int kTilePts = KIND == 1 ? 12 : (KIND == 3 ? 32 : 128);
    constexpr size_t kImgBase = KIND == 2 ? kMlpOffImg : kRevOffImg;
    // KIND 0, 4: column-split epilogue -- the four warps of a tensor-memory quadrant share every 16-feature slice (four accumulator
    // columns each) and publish it K-step by K-step, instead of owning four whole slices each and publishing rounds of four
    constexpr bool kFine = KIND == 0 || KIND == 4;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;   // same offset in both CTAs of a pair
    unsigned char* sptr = smem_dyn + (sbase - smem_u32(smem_dyn));

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const char* tc = packed + h.off_tc;     // inverse scales live in the forward engine's section
    const char* rv = packed + h.off_rev;

    const int64_t num_tiles = (n + kTilePts - 1) / kTilePts;
    const int64_t num_units = (num_tiles + CTAS - 1) / CTAS;               // work unit = one tile per CTA of the pair
    const int64_t unit0 = blockIdx.x / CTAS, unit_stride = gridDim.x / CTAS;
    const int my_units = (int)((num_units - unit0 + unit_stride - 1) / unit_stride);

    const uint32_t bars = sbase + C::kSmemBars;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

    // ---- one-time setup --------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(bar(C::kBarWFull + s), (CTAS == 2 && leader) ? 2 : 1);   // own producer (+ the peer's forward)
            mbar_init(bar(C::kBarWEmpty + s), 1);
        }
        for (int r = 0; r < 16; ++r) mbar_init(bar(C::kBarAReady + r), (kFine ? kEpiWarps / 2 : kEpiWarps) * CTAS);
        mbar_init(bar(C::kBarDFull), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if constexpr (KIND == 2) {
        float4* s_w0 = reinterpret_cast<float4*>(sptr + C::kSmemW0);
        const float4* g_w0 = reinterpret_cast<const float4*>(packed + h.off_w0);   // [256][4]: all four inputs are used
        for (int i = tid; i < 256; i += kThreads) s_w0[i] = g_w0[i];
        float* s_bias = reinterpret_cast<float*>(sptr + C::kSmemBias);
        const float* g_b0 = reinterpret_cast<const float*>(packed + h.off_b0);
        const float* g_b1 = reinterpret_cast<const float*>(packed + h.off_b);
        for (int i = tid; i < 256; i += kThreads) { s_bias[i] = g_b0[i]; s_bias[256 + i] = g_b1[i]; }
        float4* s_w2 = reinterpret_cast<float4*>(sptr + C::kSmemMlpW2);
        const float4* g_w2 = reinterpret_cast<const float4*>(rv + kMlpOffW2);
        for (int i = tid; i < 256 * 4; i += kThreads) s_w2[i] = g_w2[i];
        float* s_misc = reinterpret_cast<float*>(sptr + C::kSmemMisc);
        if (tid == 0) s_misc[0] = reinterpret_cast<const float*>(rv + kMlpOffMisc)[0];             // 1/S of the hidden layer
    } else {
        float4* s_w0 = reinterpret_cast<float4*>(sptr + C::kSmemW0);
        const float* g_w0 = reinterpret_cast<const float*>(packed + h.off_w0);
        const float* g_b0 = reinterpret_cast<const float*>(packed + h.off_b0);
        for (int i = tid; i < 256; i += kThreads) s_w0[i] = make_float4(g_w0[4 * i], g_w0[4 * i + 1], g_w0[4 * i + 2], g_b0[i]);
        float* s_bias = reinterpret_cast<float*>(sptr + C::kSmemBias);
        const float* g_bias = reinterpret_cast<const float*>(packed + h.off_b);   // f32 [3][256], omega folded
        for (int i = tid; i < 3 * 256; i += kThreads) s_bias[i] = g_bias[i];
        float2* s_w4 = reinterpret_cast<float2*>(sptr + C::kSmemW4);
        const float2* g_w4 = reinterpret_cast<const float2*>(rv + kRevOffW4);
        for (int i = tid; i < 256; i += kThreads) s_w4[i] = g_w4[i];
        float* s_misc = reinterpret_cast<float*>(sptr + C::kSmemMisc);
        if (tid < 3) s_misc[tid] = reinterpret_cast<const float*>(tc)[tid];                      // 1/S of layers 1..3
        if (tid == 3) s_misc[3] = reinterpret_cast<const float*>(packed + h.off_bout)[0];          // b4
        if (tid == 4) s_misc[4] = reinterpret_cast<const float*>(rv + kRevOffMisc)[1];             // 1/G
    }
    if (CTAS == 2) cluster_sync_all();   // both CTAs resident, barriers initialised, before any cross-CTA traffic
    if (warp == kMmaWarp) {
        if (CTAS == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(sbase + C::kSmemTmemPtr), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(sbase + C::kSmemTmemPtr), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sptr + C::kSmemTmemPtr);
    const uint32_t regionP = tmem, regionQ = tmem + 256;

    if (warp == kProducerWarp) {
        // ================= weight producer (every CTA streams its own share of the B rows) =========================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < my_units; ++it) {
                for (int g = 0; g < kGemms; ++g) {
                    const char* img = rv + kImgBase + (size_t)g * kGemmImgBytes;
                    for (int r = 0; r < 4; ++r) {
                        mbar_wait(bar(C::kBarWEmpty + stage), phase ^ 1u);
                        const uint32_t dst = sbase + C::kSmemRing + stage * C::kStageBytes;
                        const uint32_t bytes = (uint32_t)(kStageK * kPerKBytes);
                        mbar_expect_tx(bar(C::kBarWFull + stage), C::kStageBytes);
                        if (CTAS == 2) {
                            bulk_g2s(dst, img + (size_t)rank * kHalfImgBytes + (size_t)r * bytes, bytes, bar(C::kBarWFull + stage));
                        } else {
                            bulk_g2s(dst, img + (size_t)r * bytes, bytes, bar(C::kBarWFull + stage));
                            bulk_g2s(dst + bytes, img + kHalfImgBytes + (size_t)r * bytes, bytes, bar(C::kBarWFull + stage));
                        }
                        if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        if (leader) {
            // ================= MMA issuer (leader CTA only): the WHOLE warp runs the loop convergently, one elected lane issues ====
            constexpr uint32_t idesc = CTAS == 2 ? make_idesc_mn(256, 256) : make_idesc_mn(128, 128);
            uint32_t stage = 0, wphase = 0, aphase = 0;
            unsigned long long wait_a = 0, wait_w = 0;
            const long long t_begin = clock64();
            for (int it = 0; it < my_units; ++it) {
                for (int g = 0; g < kGemms; ++g) {
                    // KIND 1 runs an odd number of GEMMs per unit and fuses the unit boundary: the two halves of tensor memory swap
                    // roles from unit to unit (the next tile's layer 0 overwrites the accumulator slices of GEMM 2 as they are read)
                    const int par = (g + ((KIND == 1 || KIND == 2 || KIND == 4) ? it : 0)) & 1;
                    const uint32_t regA = par ? regionQ : regionP;
                    const uint32_t regD = par ? regionP : regionQ;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        long long t0 = dbg ? clock64() : 0;
                        mbar_wait<true>(bar(C::kBarAReady + (kFine ? kStageK * r : r)), aphase);
                        long long t1 = dbg ? clock64() : 0;
                        mbar_wait<true>(bar(C::kBarWFull + stage), wphase);
                        if (dbg) {
                            wait_a += (unsigned long long)(t1 - t0);
                            wait_w += (unsigned long long)(clock64() - t1);
                            if (blockIdx.x == 0 && lane == 0) dbg[32 + g * 4 + r] += (unsigned long long)(t1 - t0);
                        }
                        tc_fence_after();
                        const uint32_t ring = sbase + C::kSmemRing + stage * C::kStageBytes;
                        // descriptor of the stage's first K-step; the others differ in the 14-bit start-address field only (shared
                        // memory addresses stay below 2^18, so adding (byte offset >> 4) never carries out of the field)
                        const uint64_t desc0 = make_desc(ring, 2048, 128);
#pragma unroll
                        for (int tt = 0; tt < kStageK; ++tt) {
                            const int t = kStageK * r + tt;
                            if (kFine && tt > 0) {
                                const long long t2 = dbg ? clock64() : 0;
                                mbar_wait<true>(bar(C::kBarAReady + t), aphase);
                                tc_fence_after();
                                if (dbg) {
                                    wait_a += (unsigned long long)(clock64() - t2);
                                    if (blockIdx.x == 0 && lane == 0) dbg[32 + g * 4 + r] += (unsigned long long)(clock64() - t2);
                                }
                            }
                            if (dbg && blockIdx.x == 0 && it == 4 && lane == 0) dbg[64 + g * 16 + t] = (unsigned long long)clock64();   // K-step t issued
                            const uint32_t a_hi = regA + 16u * (uint32_t)t;
                            const uint32_t acc = t > 0 ? 1u : 0u;
                            if (CTAS == 2) {
                                const uint64_t bh = desc0 + (uint64_t)((uint32_t)tt * (uint32_t)(kPerKBytes >> 4)), bl = bh + (uint64_t)(4096u >> 4);
                                mma_ts_pair_elect(regD, a_hi, bh, idesc, acc);
                                mma_ts_pair_elect(regD, a_hi + 8u, bh, idesc, 1u);
                                mma_ts_pair_elect(regD, a_hi, bl, idesc, 1u);
                            } else {
#pragma unroll
                                for (int hf = 0; hf < 2; ++hf) {
                                    const uint32_t sb = ring + (uint32_t)hf * (uint32_t)(kStageK * kPerKBytes) +
                                                        (uint32_t)tt * (uint32_t)kPerKBytes;
                                    const uint64_t bh = make_desc(sb, 2048, 128), bl = make_desc(sb + 4096u, 2048, 128);
                                    const uint32_t d = regD + 128u * (uint32_t)hf;
                                    mma_ts_elect(d, a_hi, bh, idesc, acc);
                                    mma_ts_elect(d, a_hi + 8u, bh, idesc, 1u);
                                    mma_ts_elect(d, a_hi, bl, idesc, 1u);
                                }
                            }
                        }
                        if (CTAS == 2) tc_commit_pair_elect(bar(C::kBarWEmpty + stage), 3); else tc_commit_elect(bar(C::kBarWEmpty + stage));
                        if (++stage == C::kStages) { stage = 0; wphase ^= 1u; }
                    }
                    if (CTAS == 2) tc_commit_pair_elect(bar(C::kBarDFull), 3); else tc_commit_elect(bar(C::kBarDFull));
                    aphase ^= 1u;
                }
            }
            if (dbg && blockIdx.x == 0 && lane == 0) {
                dbg[0] = (unsigned long long)(clock64() - t_begin);
                dbg[1] = wait_a;
                dbg[2] = wait_w;
                dbg[3] = (unsigned long long)my_units;
            }
        } else if (lane == 0 && CTAS == 2) {
            // ================= peer CTA: forward "my half of the stage has landed" to the leader =====================
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < my_units; ++it) {
                for (int gr = 0; gr < 4 * kGemms; ++gr) {
                    mbar_wait(bar(C::kBarWFull + stage), phase);
                    mbar_arrive_cluster(mapa_rank(bar(C::kBarWFull + stage), 0));
                    if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ================= epilogue warps 0..15 ========================================================================
        const int quad = warp & 3;           // TMEM lane quadrant this warp may access
        const int grp = (warp - kFirstEpiWarp) >> 2;   // slices grp, grp+4, grp+8, grp+12 (one MMA K-step each)
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const int row = quad * 32 + lane;    // TMEM lane = point of the tile
        const float4* s_w0 = reinterpret_cast<const float4*>(sptr + C::kSmemW0);
        const float* s_bias = reinterpret_cast<const float*>(sptr + C::kSmemBias);
        const float2* s_w4 = reinterpret_cast<const float2*>(sptr + C::kSmemW4);
        const float* s_misc = reinterpret_cast<const float*>(sptr + C::kSmemMisc);
        float* s_part = reinterpret_cast<float*>(sptr + C::kSmemPart);
        float4* my_stash = stash + (size_t)blockIdx.x * (kStashBytesPerCta / 16) + row;
        const unsigned long long stash_keep = l2_keep_policy();
        uint32_t a_ready_addr[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a_ready_addr[r] = CTAS == 2 ? mapa_rank(bar(C::kBarAReady + r), 0) : bar(C::kBarAReady + r);
        uint32_t dphase = 0;
        unsigned long long wait_d = 0;
        unsigned long long* pd = (dbg && lane == 0) ? &wait_d : nullptr;

        const uint32_t a_ready_base = a_ready_addr[0];                 // barrier of K-step t: + 8 t (one contiguous array)
        auto publish_step = [&](int t) {   // KIND 0: the columns this warp just stored of K-step t are visible to the MMA issuer
            tc_wait_st();          // .sync.aligned: the whole warp's stores have completed when any lane gets past it
            tc_fence_before();
            if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(a_ready_base + 8u * (uint32_t)t); else mbar_arrive(a_ready_base + 8u * (uint32_t)t); }
        };
        auto publish = [&](int mi) {   // the slice just stored is visible to the MMA issuer
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(a_ready_addr[mi]); else mbar_arrive(a_ready_addr[mi]); }
        };

        if constexpr (KIND == 2) {
        // ---------------- ReLU level classifier: 128 points per tile, one GEMM --------------------------------------------------
        const float4* s_w2 = reinterpret_cast<const float4*>(sptr + C::kSmemMlpW2);   // [256][4 float4] = W2^T rows of 16
        float* s_lpart = reinterpret_cast<float*>(sptr + C::kSmemMlpPart);            // [4][16][128]
        const int out_f = h.out_features;
        // The unit boundary is fused as in the field engines: the first tile's layer 0 is a prologue; every later tile's layer 0 is
        // produced inside the rounds of the previous tile's epilogue, slice by slice over the accumulator columns that were just
        // read (one GEMM per unit: the two halves of tensor memory swap roles from unit to unit), so the GEMM of tile i + 1 runs
        // under the output layer of tile i instead of after it.
        auto load_x = [&](int it, int64_t& pt, bool& pvalid, float4& xin) {
            const int64_t tile = (unit0 + (int64_t)it * unit_stride) * CTAS + rank;
            pt = tile * 128 + row;
            pvalid = it < my_units && pt < n;
            xin = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pvalid) xin = reinterpret_cast<const float4*>(x)[pt];
        };
        auto layer0_slice = [&](int m, const float4 xin, uint32_t* o) {   // a0 = relu(W0 x + b0), FP32
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float r2[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int f = 16 * m + (i >> 1) + 8 * u;   // K-slot pair (i, i + 1) holds features (i / 2, 8 + i / 2)
                    const float4 w = s_w0[f];
                    // x W^T + b in the accumulation order of a row-major dot product (k = 0..3), then the bias
                    const float z = fmaf(w.w, xin.w, fmaf(w.z, xin.z, fmaf(w.y, xin.y, w.x * xin.x))) + s_bias[f];
                    r2[u] = fmaxf(z, 0.f);
                }
                split_pack<true>(r2[0], r2[1], o[i >> 1], o[8 + (i >> 1)]);
            }
        };
        int64_t pt;
        bool pvalid;
        float4 xin;
        load_x(0, pt, pvalid, xin);
        if (my_units > 0) {
#pragma unroll 1
            for (int mi = 0; mi < 4; ++mi) {
                const int m = grp + 4 * mi;
                uint32_t o[16];
                layer0_slice(m, xin, o);
                TC_ST16(regionP + lane_off + 16u * (uint32_t)m, o);
                publish(mi);
            }
        }
        for (int it = 0; it < my_units; ++it) {
            const bool has_next = it + 1 < my_units;
            const uint32_t regD = (it & 1) ? regionP : regionQ;
            int64_t npt;
            bool nvalid;
            float4 xnext;
            load_x(it + 1, npt, nvalid, xnext);   // latency hidden behind this tile's GEMM
            float acc[kMlpMaxOut];
#pragma unroll
            for (int o = 0; o < kMlpMaxOut; ++o) acc[o] = 0.f;
            {
                const float inv = s_misc[0];
                mbar_wait_timed(bar(C::kBarDFull), dphase, pd);   // every lane polls (one polling lane + __syncwarp wakes the warp later)
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
                    if (mi < 3) TC_LD16(addr + 64u, v);
                    if (has_next) {   // these accumulator columns are dead now: the next tile's layer 0 goes there first
                        uint32_t o[16];
                        layer0_slice(m, xnext, o);
                        TC_ST16(addr, o);
                        publish(mi);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int f = 16 * m + i;
                        const float a = fmaxf(fmaf(cur[i], inv, s_bias[256 + f]), 0.f);
#pragma unroll
                        for (int q = 0; q < kMlpMaxOut / 4; ++q) {
                            if (4 * q < out_f) {   // uniform
                                const float4 w = s_w2[4 * f + q];
                                acc[4 * q] = fmaf(a, w.x, acc[4 * q]);
                                acc[4 * q + 1] = fmaf(a, w.y, acc[4 * q + 1]);
                                acc[4 * q + 2] = fmaf(a, w.z, acc[4 * q + 2]);
                                acc[4 * q + 3] = fmaf(a, w.w, acc[4 * q + 3]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < kMlpMaxOut; ++o) s_lpart[(grp * kMlpMaxOut + o) * 128 + row] = acc[o];
            asm volatile("bar.sync 1, 512;" ::: "memory");
            // 512 threads write the 128 x C logits: thread (row, grp) sums the partials of outputs grp, grp + 4, ...
            if (pvalid) {
                const float* b2 = reinterpret_cast<const float*>(rv + kMlpOffMisc + 128);
                for (int o = grp; o < out_f; o += 4) {
                    const float val = ((s_lpart[(0 * kMlpMaxOut + o) * 128 + row] + s_lpart[(1 * kMlpMaxOut + o) * 128 + row]) +
                                       s_lpart[(2 * kMlpMaxOut + o) * 128 + row]) + s_lpart[(3 * kMlpMaxOut + o) * 128 + row];
                    y[pt * out_f + o] = val + b2[o];
                }
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");   // the partials are rewritten by the next tile's epilogue
            pt = npt; pvalid = nvalid; xin = xnext;
        }
        } else if constexpr (KIND == 1) {
        // ---------------- second-order forward jet: 3 points x 10 rows per quadrant ------------------------------------------
        const int pl = lane / 10;                       // point of the quadrant (3 = idle lanes 30, 31)
        const int c = lane - 10 * pl;                   // jet row: 0 value, 1..3 d/dx_i, 4..9 d2/dx_i dx_j
        const bool live = lane < 30;
        int ii = 0, jj = 0;                             // derivative directions of this row
        if (c >= 1 && c <= 3) { ii = c - 1; jj = c - 1; }
        else if (c >= 4) { const int q = c - 4; ii = q < 3 ? 0 : (q < 5 ? 1 : 2); jj = q < 3 ? q : (q < 5 ? q - 2 : 2); }
        const int src_z = 10 * pl, src_i = 10 * pl + 1 + ii, src_j = 10 * pl + 1 + jj;   // lanes holding z, dz_i, dz_j
        constexpr float kR2 = 0.0625f;                  // second-order rows are carried scaled by 2^-4 (FP16 range)
        const float sel_v = (live && c == 0) ? 1.f : 0.f, sel_1 = (live && c >= 1 && c <= 3) ? 1.f : 0.f,
                    sel_2 = (live && c >= 4) ? 1.f : 0.f;
        const float* s_w0f = reinterpret_cast<const float*>(sptr + C::kSmemW0);
        // layer 0 (FP32) of one 16-feature slice of this lane's jet row: a = sin z, da_i = cos z w_i, d2a_ij = -sin z w_i w_j
        auto layer0_jet_slice = [&](int m, float qx, float qy, float qz, uint32_t* o) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float r2[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int f = 16 * m + (i >> 1) + 8 * u;   // K-slot pair (i, i + 1) holds features (i / 2, 8 + i / 2)
                    const float4 w = s_w0[f];
                    const float z0 = fmaf(w.z, qz, fmaf(w.y, qy, fmaf(w.x, qx, w.w)));
                    const float sn = sin_layer0(z0), cn = cos_mufu(z0);
                    const float wi = s_w0f[4 * f + ii], wj = s_w0f[4 * f + jj];
                    r2[u] = sel_v * sn + sel_1 * cn * wi - (sel_2 * kR2) * sn * wi * wj;
                }
                split_pack<true>(r2[0], r2[1], o[i >> 1], o[8 + (i >> 1)]);
            }
        };
        auto load_jet_point = [&](int it, int64_t& pt, bool& pvalid, float& px, float& py, float& pz) {
            const int64_t tile = (unit0 + (int64_t)it * unit_stride) * CTAS + rank;
            pt = tile * kTilePts + 3 * quad + pl;
            pvalid = live && it < my_units && pt < n;
            px = py = pz = 0.f;
            if (pvalid) { px = x[3 * pt]; py = x[3 * pt + 1]; pz = x[3 * pt + 2]; }
        };
        // The unit boundary is fused as in KIND 0: the first tile's layer 0 is a prologue; every later tile's layer 0 is produced
        // inside the rounds of the previous tile's LAST epilogue (GEMM 2), slice by slice over the accumulator columns that were
        // just read -- with three GEMMs per unit that makes the two tensor-memory halves swap roles from unit to unit.
        int64_t pt;
        bool pvalid;
        float px, py, pz;
        load_jet_point(0, pt, pvalid, px, py, pz);
        if (my_units > 0) {
#pragma unroll 1
            for (int mi = 0; mi < 4; ++mi) {
                const int m = grp + 4 * mi;
                uint32_t o[16];
                layer0_jet_slice(m, px, py, pz, o);
                TC_ST16(regionP + lane_off + 16u * (uint32_t)m, o);
                publish(mi);
            }
        }
        for (int it = 0; it < my_units; ++it) {
            const bool has_next = it + 1 < my_units;
            int64_t npt;
            bool nvalid;
            float nx, ny, nz;
            load_jet_point(it + 1, npt, nvalid, nx, ny, nz);   // consumed three GEMMs later
            float acc = 0.f;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                const uint32_t regD = ((g + it) & 1) ? regionP : regionQ;
                const float inv = s_misc[g];
                // out = [value] sin z  +  [first, second] cos z * (own / S)  -  [second] 2^-4 sin z (dz_i / S)(dz_j / S)
                const float kb = (sel_1 + sel_2) * inv, kc = -(sel_2 * kR2) * inv * inv;
                mbar_wait_timed(bar(C::kBarDFull), dphase, pd);   // every lane polls (one polling lane + __syncwarp wakes the warp later)
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
                    if (mi < 3) TC_LD16(addr + 64u, v);
                    const float4* bp = reinterpret_cast<const float4*>(s_bias + g * 256 + 16 * m);
                    float outv[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 b = bp[q];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = 4 * q + u;
                            const float own = cur[i];
                            const float dv = __shfl_sync(0xffffffffu, own, src_z);
                            const float di = __shfl_sync(0xffffffffu, own, src_i);
                            const float dj = __shfl_sync(0xffffffffu, own, src_j);
                            const float z = fmaf(dv, inv, u == 0 ? b.x : (u == 1 ? b.y : (u == 2 ? b.z : b.w)));
                            const float sn = sin_mufu(z), cn = cos_mufu(z);
                            outv[i] = fmaf(sel_v, sn, fmaf(kb * cn, own, (kc * sn) * (di * dj)));
                        }
                    }
                    if (g < 2) {
                        uint32_t o[16];
#pragma unroll
                        for (int i = 0; i < 16; i += 2) split_pack<true>(outv[i >> 1], outv[8 + (i >> 1)], o[i >> 1], o[8 + (i >> 1)]);
                        TC_ST16(addr, o);
                        publish(mi);
                    } else {
                        // these accumulator columns are dead now (cur holds them): the next tile's layer 0 goes there first
                        if (has_next) {
                            uint32_t o[16];
                            layer0_jet_slice(m, nx, ny, nz, o);
                            TC_ST16(addr, o);
                            publish(mi);
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) acc = fmaf(outv[i], s_w4[16 * m + i].x, acc);   // linear output layer (FP32)
                    }
                }
            }
            // ---- outputs: sum the four feature-quarter partials of every jet row in a fixed order -----------------------
            s_part[grp * 128 + row] = acc;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (grp == 0 && pvalid) {
                const float val = ((s_part[row] + s_part[128 + row]) + s_part[256 + row]) + s_part[384 + row];
                if (c == 0) y[pt] = val + s_misc[3];
                else if (c <= 3) J[3 * pt + (c - 1)] = val;
                else {
                    const float hv = val * (1.f / kR2);
                    Hout[9 * pt + 3 * ii + jj] = hv;
                    Hout[9 * pt + 3 * jj + ii] = hv;
                }
            }
            // s_part is rewritten only after three more d_full phases, which every epilogue warp (the readers above included)
            // has to pass first
            pt = npt; pvalid = nvalid; px = nx; py = ny; pz = nz;
        }
        } else if constexpr (KIND == 4) {
        // ---------------- value only: the forward sweep with the column-split epilogue of KIND 0 ---------------------------------
        const int ca = 4 * (grp & 1), cb = 8 + ca, tp = grp >> 1;   // this warp's columns ca .., cb .. of the slices t = 2 j + tp
        auto load_point = [&](int it, int64_t& pt, bool& pvalid, float& px, float& py, float& pz) {
            const int64_t tile = (unit0 + (int64_t)it * unit_stride) * CTAS + rank;
            pt = tile * 128 + row;
            pvalid = it < my_units && pt < n;
            px = py = pz = 0.f;
            if (pvalid) { px = x[3 * pt]; py = x[3 * pt + 1]; pz = x[3 * pt + 2]; }
        };
        auto layer0_cols = [&](int t, float qx, float qy, float qz, uint32_t* o) {   // a0 = sin(W0' x + b0'), FP32
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float r2[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float4 w = s_w0[16 * t + (u ? cb : ca) + i];
                    r2[u] = sin_layer0(fmaf(w.z, qz, fmaf(w.y, qy, fmaf(w.x, qx, w.w))));
                }
                split_pack<true>(r2[0], r2[1], o[i], o[4 + i]);
            }
        };
        int64_t pt;
        bool pvalid;
        float px, py, pz;
        load_point(0, pt, pvalid, px, py, pz);
        if (my_units > 0) {
#pragma unroll 1
            for (int j = 0; j < 8; ++j) {
                const int t = 2 * j + tp;
                uint32_t o[8];
                layer0_cols(t, px, py, pz, o);
                const uint32_t addr = regionP + lane_off + 16u * (uint32_t)t;
                TC_ST4(addr + (uint32_t)ca, o);
                TC_ST4(addr + (uint32_t)cb, (o + 4));
                publish_step(t);
            }
        }
        for (int it = 0; it < my_units; ++it) {
            float* s_next = reinterpret_cast<float*>(sptr + C::kSmemNext) + 3 * (grp * 128 + row);
            const bool has_next = it + 1 < my_units;
            float acc_y = 0.f;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                const uint32_t regD = ((g + it) & 1) ? regionP : regionQ;
                const float inv = s_misc[g];
                if (g == 0) {   // the next tile's point: consumed two GEMMs later
                    int64_t npt; bool nvalid; float nx, ny, nz;
                    load_point(it + 1, npt, nvalid, nx, ny, nz);
                    s_next[0] = nx; s_next[1] = ny; s_next[2] = nz;
                }
                mbar_wait_timed(bar(C::kBarDFull), dphase, pd);   // every lane polls (one polling lane + __syncwarp wakes the warp later)
                dphase ^= 1u;
                tc_fence_after();
                uint32_t v[8];
                TC_LD4(regD + lane_off + 16u * (uint32_t)tp + (uint32_t)ca, v);
                TC_LD4(regD + lane_off + 16u * (uint32_t)tp + (uint32_t)cb, (v + 4));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int t = 2 * j + tp;
                    const uint32_t addr = regD + lane_off + 16u * (uint32_t)t;
                    tc_wait_ld();
                    float cur[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) cur[i] = __uint_as_float(v[i]);
                    if (j < 7) {
                        TC_LD4(addr + 32u + (uint32_t)ca, v);
                        TC_LD4(addr + 32u + (uint32_t)cb, (v + 4));
                    }
                    const float4 ba = *reinterpret_cast<const float4*>(s_bias + g * 256 + 16 * t + ca);
                    const float4 bb = *reinterpret_cast<const float4*>(s_bias + g * 256 + 16 * t + cb);
                    const float zz[8] = {fmaf(cur[0], inv, ba.x), fmaf(cur[1], inv, ba.y), fmaf(cur[2], inv, ba.z), fmaf(cur[3], inv, ba.w),
                                         fmaf(cur[4], inv, bb.x), fmaf(cur[5], inv, bb.y), fmaf(cur[6], inv, bb.z), fmaf(cur[7], inv, bb.w)};
                    uint32_t o[8];
                    if (g < 2) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) split_pack<true>(sin_mufu(zz[i]), sin_mufu(zz[4 + i]), o[i], o[4 + i]);
                        TC_ST4(addr + (uint32_t)ca, o);
                        TC_ST4(addr + (uint32_t)cb, (o + 4));
                        publish_step(t);
                    } else {
                        // these accumulator columns are dead now: the next tile's layer 0 goes there first, the value is summed behind it
                        if (has_next) {
                            layer0_cols(t, s_next[0], s_next[1], s_next[2], o);
                            TC_ST4(addr + (uint32_t)ca, o);
                            TC_ST4(addr + (uint32_t)cb, (o + 4));
                            publish_step(t);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {   // y += w4 . sin z3
                            acc_y = fmaf(sin_mufu(zz[i]), s_w4[16 * t + ca + i].x, acc_y);
                            acc_y = fmaf(sin_mufu(zz[4 + i]), s_w4[16 * t + cb + i].x, acc_y);
                        }
                    }
                }
            }
            // ---- output: sum the four partials of every point in a fixed order ---------------------------------------------
            s_part[grp * 128 + row] = acc_y;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (grp == 0 && pvalid)
                y[pt] = (((s_part[row] + s_part[128 + row]) + s_part[256 + row]) + s_part[384 + row]) + s_misc[3];
            // s_part is rewritten only after three more d_full phases, which every epilogue warp (the readers above included)
            // has to pass first
            {
                const int64_t tile = (unit0 + (int64_t)(it + 1) * unit_stride) * CTAS + rank;
                pt = tile * 128 + row;
                pvalid = has_next && pt < n;
                if (has_next) { px = s_next[0]; py = s_next[1]; pz = s_next[2]; }
            }
        }
        } else if constexpr (KIND == 3) {
        // ---------------- forward-over-reverse: 32 points x (primal row + three tangent rows) ------------------------------
        //   forward   z = a W'^T + b'            a = sin z                 c = cos z
        //             zt_i = at_i W'^T           at_i = c (.) zt_i         ct_i = -a (.) zt_i                 (tangents d/dx_i)
        //   backward  g_{l-1}    = (g_l W') (.) c_{l-1}
        //             gt_{l-1,i} = (gt_{l,i} W') (.) c_{l-1} + (g_l W') (.) ct_{l-1,i}
        //   input     dy/dx = g_0 W_0' ,  H[i,:] = gt_{0,i} W_0'
        // Every row stashes ONE float per feature and layer for the backward sweep (primal: c, tangent: ct), so the scratch has
        // the KIND 0 layout; a tangent row fetches its point's z / c / upstream product from the primal lane by shuffle.
        const int c = lane & 3;                         // 0: primal row; 1..3: tangent along x_{c-1}
        const int src = lane & ~3;                      // lane of this point's primal row
        const int ii = c > 0 ? c - 1 : 0;
        const float selP = c == 0 ? 1.f : 0.f, selT = c == 0 ? 0.f : 1.f;
        constexpr float kT = 0.03125f;                  // tangent rows of the backward sweep are carried scaled by 2^-5 (FP16 range)
        const float kTs = selT * kT;
        const float phase0 = selT * kHalfPi;            // layer 0: one MUFU per row-feature, cos z = sin(z + pi/2) on tangent rows
        const float* s_w0f = reinterpret_cast<const float*>(sptr + C::kSmemW0);
        auto load_point = [&](int it, int64_t& pt, bool& pvalid, float& px, float& py, float& pz) {
            const int64_t tile = (unit0 + (int64_t)it * unit_stride) * CTAS + rank;
            pt = tile * kTilePts + (row >> 2);
            pvalid = it < my_units && pt < n;
            px = py = pz = 0.f;
            if (pvalid) { px = x[3 * pt]; py = x[3 * pt + 1]; pz = x[3 * pt + 2]; }
        };
        // layer 0 (FP32): a_0 = sin z_0 (primal), da_0/dx_i = cos z_0 W_0'[:, i] (tangent).  The quarter turn is added to the bias
        // BEFORE the dot product, so the argument carries the same FP32 rounding as z_0 itself and nothing more.
        auto layer0_slice = [&](int m, float qx, float qy, float qz, uint32_t* o) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float r2[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int f = 16 * m + (i >> 1) + 8 * u;   // K-slot pair (i, i + 1) holds features (i / 2, 8 + i / 2)
                    const float4 w = s_w0[f];
                    const float z0 = fmaf(w.z, qz, fmaf(w.y, qy, fmaf(w.x, qx, w.w + phase0)));
                    const float wi = s_w0f[4 * f + ii];
                    r2[u] = sin_layer0(z0) * (c == 0 ? 1.f : wi);
                }
                split_pack<true>(r2[0], r2[1], o[i >> 1], o[8 + (i >> 1)]);
            }
        };
        int64_t pt;
        bool pvalid;
        float px, py, pz;
        load_point(0, pt, pvalid, px, py, pz);
        if (my_units > 0) {
#pragma unroll 1
            for (int mi = 0; mi < 4; ++mi) {
                const int m = grp + 4 * mi;
                uint32_t o[16];
                layer0_slice(m, px, py, pz, o);
                TC_ST16(regionP + lane_off + 16u * (uint32_t)m, o);
                publish(mi);
            }
        }
        for (int it = 0; it < my_units; ++it) {
            float* s_next = reinterpret_cast<float*>(sptr + C::kSmemNext) + 3 * (grp * 128 + row);
            const bool has_next = it + 1 < my_units;
            float acc_y = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
            for (int g = 0; g < 6; ++g) {
                const uint32_t regD = (g & 1) ? regionP : regionQ;
                const float inv = s_misc[g < 3 ? g : 5 - g];              // 1/S of the layer this GEMM multiplies by
                const float invT = selT * inv;
                float4 cs[4];                                             // this row's stash of the current slice
                if (g == 3) {
                    int64_t npt; bool nvalid; float nx, ny, nz;
                    load_point(it + 1, npt, nvalid, nx, ny, nz);
                    s_next[0] = nx; s_next[1] = ny; s_next[2] = nz;
                }
                if (g == 3 || g == 4) {
                    const float4* sp = my_stash + (size_t)(((4 - g) * 16 + grp) * 4) * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j) cs[j] = ld_keep4(sp + j * 128, stash_keep);
                }
                mbar_wait_timed(bar(C::kBarDFull), dphase, pd);   // every lane polls (one polling lane + __syncwarp wakes the warp later)
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
                    if (mi < 3) TC_LD16(addr + 64u, v);
                    uint32_t o[16];
                    if (g < 3) {
                        // forward through layer g + 1
                        const float4* bp = reinterpret_cast<const float4*>(s_bias + g * 256 + 16 * m);
                        float av[16], sv[16];                             // next operand / what the backward sweep needs of this layer
                        // packed FP32 math (FFMA2 on pairs of features): 7 instead of 10 issue slots per row-feature
                        const float2 inv2 = make_float2(inv, inv), invT2 = make_float2(invT, invT), selP2 = make_float2(selP, selP);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 b = bp[q];
#pragma unroll
                            for (int u = 0; u < 4; u += 2) {
                                const int i = 4 * q + u;
                                const float2 own2 = make_float2(cur[i], cur[i + 1]);
                                const float2 zp2 = make_float2(__shfl_sync(0xffffffffu, cur[i], src), __shfl_sync(0xffffffffu, cur[i + 1], src));
                                const float2 z2 = __ffma2_rn(zp2, inv2, u == 0 ? make_float2(b.x, b.y) : make_float2(b.z, b.w));
                                const float2 sn2 = make_float2(sin_mufu(z2.x), sin_mufu(z2.y)), cn2 = make_float2(cos_mufu(z2.x), cos_mufu(z2.y));
                                const float2 zt2 = __fmul2_rn(own2, invT2);       // tangent rows: dz/dx_i; 0 on the primal row
                                const float2 a2 = __ffma2_rn(cn2, zt2, __fmul2_rn(selP2, sn2));                            // sin z | cos z dz/dx_i
                                const float2 s2 = __ffma2_rn(make_float2(-sn2.x, -sn2.y), zt2, __fmul2_rn(selP2, cn2));    // cos z | -sin z dz/dx_i
                                av[i] = a2.x; av[i + 1] = a2.y;
                                sv[i] = s2.x; sv[i + 1] = s2.y;
                            }
                        }
                        if (g < 2) {
#pragma unroll
                            for (int i = 0; i < 16; i += 2) split_pack<true>(av[i >> 1], av[8 + (i >> 1)], o[i >> 1], o[8 + (i >> 1)]);
                            TC_ST16(addr, o);
                            publish(mi);
                            float4* sp = my_stash + (size_t)((g * 16 + m) * 4) * 128;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                st_keep4(sp + j * 128, make_float4(sv[4 * j], sv[4 * j + 1], sv[4 * j + 2], sv[4 * j + 3]), stash_keep);
                        } else {
                            // last hidden layer: the backward sweep starts with G w4 (.) c_3 (primal) and kT G w4 (.) ct_3 (tangents)
                            const float ks = selP + kTs;
#pragma unroll
                            for (int i = 0; i < 16; i += 2) {
                                const float2 w0 = s_w4[16 * m + (i >> 1)], w1 = s_w4[16 * m + 8 + (i >> 1)];
                                split_pack<true>(sv[i >> 1] * (ks * w0.y), sv[8 + (i >> 1)] * (ks * w1.y), o[i >> 1], o[8 + (i >> 1)]);
                            }
                            TC_ST16(addr, o);
                            publish(mi);
#pragma unroll
                            for (int i = 0; i < 16; ++i) acc_y = fmaf(av[i], s_w4[16 * m + i].x, acc_y);   // y on the primal row
                        }
                    } else if (g < 5) {
                        // backward through layer 6 - g
                        float gg[16];
                        const float2 inv2 = make_float2(inv, inv), kTs2 = make_float2(kTs, kTs);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
#pragma unroll
                            for (int u = 0; u < 4; u += 2) {
                                const int i = 4 * j + u;
                                const float2 st2 = u == 0 ? make_float2(cs[j].x, cs[j].y) : make_float2(cs[j].z, cs[j].w);
                                const float2 own2 = __fmul2_rn(make_float2(cur[i], cur[i + 1]), inv2);
                                // the point's cosine and the point's upstream product, from the primal lane
                                const float2 pc2 = make_float2(__shfl_sync(0xffffffffu, st2.x, src), __shfl_sync(0xffffffffu, st2.y, src));
                                const float2 pD2 = make_float2(__shfl_sync(0xffffffffu, own2.x, src), __shfl_sync(0xffffffffu, own2.y, src));
                                const float2 g2 = __ffma2_rn(__fmul2_rn(kTs2, pD2), st2, __fmul2_rn(own2, pc2));
                                gg[i] = g2.x; gg[i + 1] = g2.y;
                            }
                        }
                        if (mi < 3) {
                            const float4* sp = my_stash + (size_t)(((4 - g) * 16 + m + 4) * 4) * 128;
#pragma unroll
                            for (int j = 0; j < 4; ++j) cs[j] = ld_keep4(sp + j * 128, stash_keep);
                        }
#pragma unroll
                        for (int i = 0; i < 16; i += 2) split_pack<true>(gg[i >> 1], gg[8 + (i >> 1)], o[i >> 1], o[8 + (i >> 1)]);
                        TC_ST16(addr, o);
                        publish(mi);
                    } else {
                        if (has_next) {
                            layer0_slice(m, s_next[0], s_next[1], s_next[2], o);
                            TC_ST16(addr, o);
                            publish(mi);
                        }
                        // input layer: g_0 = (D / S) (.) c_0 [+ kT (D_primal / S) (.) ct_0], c_0 / ct_0 recomputed from the point;
                        // the primal lane hands -kT (D / S) sin z_0 to its tangents
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int f = 16 * m + i;
                            const float4 w = s_w0[f];
                            const float z0 = fmaf(w.z, pz, fmaf(w.y, py, fmaf(w.x, px, w.w)));
                            const float own = cur[i] * inv;
                            const float q = __shfl_sync(0xffffffffu, (own * -kT) * sin_layer0(z0), src);
                            const float g0 = fmaf(selT * q, s_w0f[4 * f + ii], own * cos_mufu(z0));
                            gx = fmaf(g0, w.x, gx);
                            gy = fmaf(g0, w.y, gy);
                            gz = fmaf(g0, w.z, gz);
                        }
                    }
                }
            }

            // ---- outputs: sum the four feature-quarter partials of every row in a fixed order ----------------------------
            s_part[(grp * 4 + 0) * 128 + row] = acc_y;
            s_part[(grp * 4 + 1) * 128 + row] = gx;
            s_part[(grp * 4 + 2) * 128 + row] = gy;
            s_part[(grp * 4 + 3) * 128 + row] = gz;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (grp == 0) {
                float r4[4];
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    r4[o] = ((s_part[(0 * 4 + o) * 128 + row] + s_part[(1 * 4 + o) * 128 + row]) +
                             s_part[(2 * 4 + o) * 128 + row]) + s_part[(3 * 4 + o) * 128 + row];
                const float inv_g = s_misc[4];
                // Hessian rows come out of three separate tangent sweeps: symmetrise, H[i][k] = (h_i[k] + h_k[i]) / 2
                const float hs = inv_g * (1.f / kT);
                const float hrow[3] = {r4[1] * hs, r4[2] * hs, r4[3] * hs};
                float other[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float t0 = __shfl_sync(0xffffffffu, hrow[0], src + 1 + k), t1 = __shfl_sync(0xffffffffu, hrow[1], src + 1 + k),
                                t2 = __shfl_sync(0xffffffffu, hrow[2], src + 1 + k);
                    other[k] = ii == 0 ? t0 : (ii == 1 ? t1 : t2);
                }
                if (pvalid) {
                    if (c == 0) {
                        y[pt] = r4[0] + s_misc[3];
                        J[3 * pt] = r4[1] * inv_g;
                        J[3 * pt + 1] = r4[2] * inv_g;
                        J[3 * pt + 2] = r4[3] * inv_g;
                    } else {
#pragma unroll
                        for (int k = 0; k < 3; ++k) Hout[9 * pt + 3 * ii + k] = 0.5f * (hrow[k] + other[k]);
                    }
                }
            }
            {   // the parked point becomes the current one (hazards into the next unit: as in KIND 0)
                const int64_t tile = (unit0 + (int64_t)(it + 1) * unit_stride) * CTAS + rank;
                pt = tile * kTilePts + (row >> 2);
                pvalid = has_next && pt < n;
                if (has_next) { px = s_next[0]; py = s_next[1]; pz = s_next[2]; }
            }
        }
        } else {
        // The unit boundary is fused: the first tile's layer 0 is a prologue, every later tile's layer 0 is produced INSIDE
        // the rounds of the previous tile's last epilogue (GEMM 5), slice by slice over the accumulator columns that were
        // just read -- so the MMAs of the next tile's first GEMM start after one epilogue round, like any other GEMM,
        // instead of after the whole input-layer epilogue, the output reduction and a fresh layer 0.
        auto load_point = [&](int it, int64_t& pt, bool& pvalid, float& px, float& py, float& pz) {
            const int64_t tile = (unit0 + (int64_t)it * unit_stride) * CTAS + rank;
            pt = tile * 128 + row;
            pvalid = it < my_units && pt < n;
            px = py = pz = 0.f;
            if (pvalid) { px = x[3 * pt]; py = x[3 * pt + 1]; pz = x[3 * pt + 2]; }
        };
        // Column split: two warps share a 16-feature slice.  This warp owns accumulator columns ca .. ca + 3 and cb .. cb + 3
        // (ca = 4 (grp & 1), cb = 8 + ca) of the slices t = 2 j + (grp >> 1), j = 0 .. 7 -- with the K-slot order of the weight images
        // those are exactly the columns its FP16 hi words (ca ..) and lo words (cb ..) of K-step t go to.  A slice is complete, and
        // its K-step published, after half of the math a whole-slice owner needs, two slices at a time: the first MMAs of the
        // next GEMM start that much earlier (the hand-over is the engine's largest idle term), and the MMAs of the last K-steps are
        // issued as soon as THEIR columns are done instead of after a round of four.
        const int ca = 4 * (grp & 1), cb = 8 + ca, tp = grp >> 1;
        auto layer0_cols = [&](int t, float qx, float qy, float qz, uint32_t* o) {   // a0 = sin(W0' x + b0'), FP32
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float r2[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float4 w = s_w0[16 * t + (u ? cb : ca) + i];
                    r2[u] = sin_layer0(fmaf(w.z, qz, fmaf(w.y, qy, fmaf(w.x, qx, w.w))));
                }
                split_pack<true>(r2[0], r2[1], o[i], o[4 + i]);    // hi word ca + i, lo word cb + i
            }
        };
        int64_t pt;
        bool pvalid;
        float px, py, pz;
        load_point(0, pt, pvalid, px, py, pz);
        if (my_units > 0) {
#pragma unroll 1
            for (int j = 0; j < 8; ++j) {
                const int t = 2 * j + tp;
                uint32_t o[8];
                layer0_cols(t, px, py, pz, o);
                const uint32_t addr = regionP + lane_off + 16u * (uint32_t)t;
                TC_ST4(addr + (uint32_t)ca, o);
                TC_ST4(addr + (uint32_t)cb, (o + 4));
                publish_step(t);
            }
        }
        for (int it = 0; it < my_units; ++it) {
            // the next tile's point is prefetched while this tile's GEMMs run and parked in this thread's own shared-memory
            // slot: three registers less across the backward sweep
            float* s_next = reinterpret_cast<float*>(sptr + C::kSmemNext) + 3 * (grp * 128 + row);
            const bool has_next = it + 1 < my_units;

            float acc_y = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
            for (int g = 0; g < 6; ++g) {
                const uint32_t regD = (g & 1) ? regionP : regionQ;
                const float inv = s_misc[g < 3 ? g : 5 - g];              // 1/S of the layer this GEMM multiplies by
                float4 csa = make_float4(0.f, 0.f, 0.f, 0.f), csb = csa;  // stashed cosines of this warp's columns of the current slice
                if (g == 3) {   // global-load latency hidden behind two GEMMs
                    int64_t npt; bool nvalid; float nx, ny, nz;
                    load_point(it + 1, npt, nvalid, nx, ny, nz);
                    s_next[0] = nx; s_next[1] = ny; s_next[2] = nz;
                }
                // stash: [layer][slice][4][128 rows] float4 -- words 2 (grp & 1) and 2 (grp & 1) + 1 of a slice hold the cosines of this
                // thread's columns ca .. and cb .., written and read by the same thread (coalesced 512 B per warp and word)
                auto stash_at = [&](int slot, int t) { return my_stash + (size_t)((slot * 16 + t) * 4 + 2 * (grp & 1)) * 128; };
                if (g == 3 || g == 4) {
                    const float4* sp = stash_at(4 - g, tp);
                    csa = ld_keep4(sp, stash_keep); csb = ld_keep4(sp + 128, stash_keep);
                }
                mbar_wait_timed(bar(C::kBarDFull), dphase, pd);   // every lane polls (one polling lane + __syncwarp wakes the warp later)
                dphase ^= 1u;
                tc_fence_after();
                const bool trace = dbg && blockIdx.x == 0 && it == 4 && warp == kFirstEpiWarp && lane == 0;
                if (trace) dbg[160 + g] = (unsigned long long)clock64();            // accumulator of GEMM g seen complete
                uint32_t v[8];
                TC_LD4(regD + lane_off + 16u * (uint32_t)tp + (uint32_t)ca, v);
                TC_LD4(regD + lane_off + 16u * (uint32_t)tp + (uint32_t)cb, (v + 4));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int t = 2 * j + tp;
                    const uint32_t addr = regD + lane_off + 16u * (uint32_t)t;
                    tc_wait_ld();
                    float cur[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) cur[i] = __uint_as_float(v[i]);    // features ca .. ca + 3, cb .. cb + 3
                    if (j < 7) {   // prefetch this warp's next slice under the math
                        TC_LD4(addr + 32u + (uint32_t)ca, v);
                        TC_LD4(addr + 32u + (uint32_t)cb, (v + 4));
                    }
                    // Front: what the next GEMM's operand needs.  A K-step is published one iteration LATE (after the next slice's
                    // math, when its tensor-memory store has long completed: no stall on tcgen05.wait::st) -- except the first
                    // one of a GEMM, which is what the idle tensor pipe is waiting for.
                    uint32_t o[8];
                    float zz[8];
                    float2 w4a[4], w4b[4];
                    const bool stores = g < 5 || has_next;
                    if (g < 3) {
                        // forward: z = D / S + b'
                        const float4 ba = *reinterpret_cast<const float4*>(s_bias + g * 256 + 16 * t + ca);
                        const float4 bb = *reinterpret_cast<const float4*>(s_bias + g * 256 + 16 * t + cb);
                        zz[0] = fmaf(cur[0], inv, ba.x); zz[1] = fmaf(cur[1], inv, ba.y); zz[2] = fmaf(cur[2], inv, ba.z); zz[3] = fmaf(cur[3], inv, ba.w);
                        zz[4] = fmaf(cur[4], inv, bb.x); zz[5] = fmaf(cur[5], inv, bb.y); zz[6] = fmaf(cur[6], inv, bb.z); zz[7] = fmaf(cur[7], inv, bb.w);
                        if (g < 2) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) split_pack<true>(sin_mufu(zz[i]), sin_mufu(zz[4 + i]), o[i], o[4 + i]);
                        } else {
                            // last hidden layer: start the backward sweep with G w4 (.) cos z3
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                w4a[i] = s_w4[16 * t + ca + i]; w4b[i] = s_w4[16 * t + cb + i];
                                split_pack<true>(cos_mufu(zz[i]) * w4a[i].y, cos_mufu(zz[4 + i]) * w4b[i].y, o[i], o[4 + i]);
                            }
                        }
                    } else if (g < 5) {
                        // backward through layer 6-g: g_{l-1} = (D / S) (.) cos z_{l-1}   (cosines from the stash)
                        const float gg[8] = {cur[0] * inv * csa.x, cur[1] * inv * csa.y, cur[2] * inv * csa.z, cur[3] * inv * csa.w,
                                             cur[4] * inv * csb.x, cur[5] * inv * csb.y, cur[6] * inv * csb.z, cur[7] * inv * csb.w};
                        if (j < 7) {
                            const float4* sp = stash_at(4 - g, t + 2);
                            csa = ld_keep4(sp, stash_keep); csb = ld_keep4(sp + 128, stash_keep);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) split_pack<true>(gg[i], gg[4 + i], o[i], o[4 + i]);
                    } else if (has_next) {
                        // these accumulator columns (region P, slice t) are dead now (cur holds them): the next tile's layer 0 goes there
                        layer0_cols(t, s_next[0], s_next[1], s_next[2], o);
                    }
                    if (stores) {
                        TC_ST4(addr + (uint32_t)ca, o);
                        TC_ST4(addr + (uint32_t)cb, (o + 4));
                        publish_step(t);
                        if (trace) dbg[168 + g * 8 + j] = (unsigned long long)clock64();   // slice of iteration j stored (published one later)
                        if (dbg && blockIdx.x == 0 && it == 4 && g == 2 && j == 0 && lane == 0) dbg[216 + warp] = (unsigned long long)clock64();
                    }
                    // Back: everything else of the slice, off the hand-over path.
                    if (g < 2) {
                        float4* sp = stash_at(g, t);
                        st_keep4(sp, make_float4(cos_mufu(zz[0]), cos_mufu(zz[1]), cos_mufu(zz[2]), cos_mufu(zz[3])), stash_keep);
                        st_keep4(sp + 128, make_float4(cos_mufu(zz[4]), cos_mufu(zz[5]), cos_mufu(zz[6]), cos_mufu(zz[7])), stash_keep);
                    } else if (g == 2) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {   // y += w4 . sin z3
                            acc_y = fmaf(sin_mufu(zz[i]), w4a[i].x, acc_y);
                            acc_y = fmaf(sin_mufu(zz[4 + i]), w4b[i].x, acc_y);
                        }
                    } else if (g == 5) {
                        // input layer: g_0 = (D / S) (.) cos z_0 (recomputed), dy/dx += g_0 W0'
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 w = s_w0[16 * t + (i < 4 ? ca + i : cb + i - 4)];
                            const float z0 = fmaf(w.z, pz, fmaf(w.y, py, fmaf(w.x, px, w.w)));
                            const float g0 = cur[i] * inv * cos_mufu(z0);
                            gx = fmaf(g0, w.x, gx);
                            gy = fmaf(g0, w.y, gy);
                            gz = fmaf(g0, w.z, gz);
                        }
                    }
                }
            }

            // ---- outputs: sum the four feature-quarter partials of every point in a fixed order -------------------
            s_part[(grp * 4 + 0) * 128 + row] = acc_y;
            s_part[(grp * 4 + 1) * 128 + row] = gx;
            s_part[(grp * 4 + 2) * 128 + row] = gy;
            s_part[(grp * 4 + 3) * 128 + row] = gz;
            asm volatile("bar.sync 1, 512;" ::: "memory");  // the 16 epilogue warps only
            if (grp == 0 && pvalid) {
                float r4[4];
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    r4[o] = ((s_part[(0 * 4 + o) * 128 + row] + s_part[(1 * 4 + o) * 128 + row]) +
                             s_part[(2 * 4 + o) * 128 + row]) + s_part[(3 * 4 + o) * 128 + row];
                const float inv_g = s_misc[4];
                if (peers.mc != nullptr) {
                    // fused all-gather, NVLink multicast: ONE store per point, the NVSwitch replicates the row into the
                    // gathered buffer of every rank (this one included)
                    float4* q = peers.mc + pt;
                    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};"
                                 ::"l"(q), "f"(r4[0] + s_misc[3]), "f"(r4[1] * inv_g), "f"(r4[2] * inv_g), "f"(r4[3] * inv_g) : "memory");
                } else if (peers.n > 0) {
                    // fused all-gather, peer-to-peer: the row goes straight into every rank's gathered buffer over NVLink
                    const float4 rowv = make_float4(r4[0] + s_misc[3], r4[1] * inv_g, r4[2] * inv_g, r4[3] * inv_g);
#pragma unroll
                    for (int g = 0; g < kMaxRowPeers; ++g)
                        if (g < peers.n) peers.dst[g][pt] = rowv;
                } else if (Hout != nullptr) {   // packed rows (value, d/dx, d/dy, d/dz): one 16-byte store per point
                    reinterpret_cast<float4*>(Hout)[pt] = make_float4(r4[0] + s_misc[3], r4[1] * inv_g, r4[2] * inv_g, r4[3] * inv_g);
                } else {
                    y[pt] = r4[0] + s_misc[3];
                    J[3 * pt] = r4[1] * inv_g;
                    J[3 * pt + 1] = r4[2] * inv_g;
                    J[3 * pt + 2] = r4[3] * inv_g;
                }
            }
            // Hazards into the next unit: the next tile's layer 0 overwrote exactly the accumulator slices this warp had
            // finished reading (tcgen05.wait::ld before the math of each round); s_part is rewritten only after six more
            // d_full phases, which need every epilogue warp (including the readers above) to have passed this point.
            {   // the parked point becomes the current one
                const int64_t tile = (unit0 + (int64_t)(it + 1) * unit_stride) * CTAS + rank;
                pt = tile * 128 + row;
                pvalid = has_next && pt < n;
                if (has_next) { px = s_next[0]; py = s_next[1]; pz = s_next[2]; }
            }
        }
        }
        if (dbg && blockIdx.x == 0 && lane == 0) dbg[8 + warp - kFirstEpiWarp] = wait_d;
    }

    // ---- teardown ----------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (CTAS == 2) cluster_sync_all();   // the peer's tensor memory is part of the leader's MMAs until both are done
    if (warp == kMmaWarp) {
        if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// ---- pack-time kernels ------------------------------------------------------------------------------------------------
// power-of-two scale that puts max |w * mul| into [top/2, top)
__device__ __forceinline__ float pow2_scale_to(float maxabs, int top_exp) {
    if (!(maxabs > 0.f) || !isfinite(maxabs)) return 1.f;
    int e;
    frexpf(maxabs, &e);
    return ldexpf(1.f, top_exp - e);
}

// One GEMM image: B[nb][kb] = S * omega0 * (transpose ? W[kb][nb] : W[nb][kb]), FP16 hi / lo, split in two N-halves,
// each K-step stored as the K-major no-swizzle core-matrix image (8 rows x 16 B) the UMMA descriptor walks
// (LBO = 2048 between the two K-halves, SBO = 128 between 8-row groups).
__global__ void rev_pack_gemm_kernel(const float* __restrict__ W, float omega0, const uint32_t* __restrict__ maxabs_bits,
                                     int transpose, char* __restrict__ img) {
    const float scale = pow2_scale_to(omega0 * __uint_as_float(*maxabs_bits), 4);   // same S as the forward engine
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 256 * 256) return;
    // K slot kb of the image: slots (2c, 2c + 1) of a 16-wide K-step hold features (c, 8 + c) of that step, so that the FP16 hi
    // word c and lo word 8 + c an epilogue thread writes over the accumulator columns cover exactly the columns they came from
    const int nb = idx >> 8, kb = idx & 255;
    const int kf = (kb & ~15) | ((kb & 1) ? 8 + ((kb & 15) >> 1) : ((kb & 15) >> 1));
    const float w = transpose ? W[kf * 256 + nb] : W[nb * 256 + kf];
    const float v = omega0 * w * scale;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const int half = nb >> 7, nl = nb & 127, t = kb >> 4, kc = (kb >> 3) & 1;
    const size_t off = (size_t)half * kHalfImgBytes + (size_t)t * kPerKBytes + (size_t)kc * 2048 + (size_t)(nl >> 3) * 128 +
                       (size_t)(nl & 7) * 16 + (size_t)(kb & 7) * 2;
    *reinterpret_cast<__half*>(img + off) = hi;
    *reinterpret_cast<__half*>(img + off + 4096) = lo;
}

// Column abs-sum norm of one hidden layer, omega folded: max_k sum_n |W'[n][k]| -- the factor by which the backward
// sweep can at most amplify max |g| through this layer.  misc[8 + l] receives it.
__global__ void rev_colnorm_kernel(const float* __restrict__ W, float omega0, float* __restrict__ out) {
    __shared__ float s_m[8];
    const int k = threadIdx.x;   // 256 threads = 256 columns
    float sum = 0.f;
    for (int nrow = 0; nrow < 256; ++nrow) sum += fabsf(omega0 * W[nrow * 256 + k]);
    for (int o = 16; o > 0; o >>= 1) sum = fmaxf(sum, __shfl_xor_sync(0xffffffffu, sum, o));
    if ((k & 31) == 0) s_m[k >> 5] = sum;
    __syncthreads();
    if (k == 0) { float m = s_m[0]; for (int i = 1; i < 8; ++i) m = fmaxf(m, s_m[i]); *out = m; }
}

__global__ void rev_pack_out_kernel(const float* __restrict__ W4, const uint32_t* __restrict__ maxabs_bits,
                                    char* __restrict__ rv) {
    // Range scale of the backward sweep.  max |G w4| in [1, 2) leaves 2^15 of FP16 headroom, which random-init and
    // typical trained fields never approach (measured max |G g| ~ 4); if the worst-case amplification of the three
    // layers (product of their column abs-sum norms) could exceed it, G is lowered by the missing power of two, so
    // the hi parts can never overflow (the lo parts then lose that many bits instead).
    float G = pow2_scale_to(__uint_as_float(*maxabs_bits), 1);
    {
        const float* norms = reinterpret_cast<const float*>(rv + kRevOffMisc) + 8;
        const float worst = 2.f * norms[0] * norms[1] * norms[2];   // bound on max |G g_0| with max |G w4| < 2
        if (worst > 32768.f && isfinite(worst)) { int e; frexpf(worst / 32768.f, &e); G = ldexpf(G, -e); }
    }
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f == 0) {
        float* misc = reinterpret_cast<float*>(rv + kRevOffMisc);
        misc[0] = G;
        misc[1] = 1.f / G;
    }
    if (f < 256) reinterpret_cast<float2*>(rv + kRevOffW4)[f] = make_float2(W4[f], G * W4[f]);
}

__global__ void mlp_maxabs_kernel(const float* __restrict__ W, int count, uint32_t* __restrict__ slot) {
    float m = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(W[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(m));
}

__global__ void mlp_pack_out_kernel(const float* __restrict__ W2, const float* __restrict__ b2, int out_f, char* __restrict__ rv) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // f * 16 + o
    const uint32_t* mx = reinterpret_cast<const uint32_t*>(rv + kMlpOffMisc + 64);
    if (idx == 0) reinterpret_cast<float*>(rv + kMlpOffMisc)[0] = 1.f / pow2_scale_to(__uint_as_float(*mx), 4);
    if (idx < kMlpMaxOut) reinterpret_cast<float*>(rv + kMlpOffMisc + 128)[idx] = idx < out_f ? b2[idx] : 0.f;
    if (idx < 256 * kMlpMaxOut) {
        const int f = idx / kMlpMaxOut, o = idx % kMlpMaxOut;
        reinterpret_cast<float*>(rv + kMlpOffW2)[idx] = o < out_f ? W2[o * 256 + f] : 0.f;
    }
}

template <int CTAS, int KIND>
int launch_rev(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, float* J, float* H, void* ws,
               cudaStream_t stream, const RowPeers peers = RowPeers{}) {
    using C = RevCfg<CTAS, KIND>;
    auto kern = siren_rev_kernel<CTAS, KIND>;
    const int smem = (int)(KIND == 2 ? C::kSmemTotalMlp : C::kSmemTotal) + 1024;
    INF3DP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    constexpr int kTilePts = KIND == 1 ? 12 : (KIND == 3 ? 32 : 128);
    const int64_t tiles = (n + kTilePts - 1) / kTilePts;
    const int64_t units = (tiles + CTAS - 1) / CTAS;
    const int64_t max_units = siren_sms() / CTAS;
    const int grid = (int)(units < max_units ? units : max_units) * CTAS;
    static const bool debug = getenv("INF3DP_TC_DEBUG") != nullptr;
    unsigned long long* dbg = nullptr;
    if (debug) {
        INF3DP_CUDA(cudaMalloc(&dbg, 256 * sizeof(unsigned long long)));
        INF3DP_CUDA(cudaMemsetAsync(dbg, 0, 256 * sizeof(unsigned long long), stream));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    INF3DP_CUDA(cudaLaunchKernelEx(&cfg, kern, (const char*)packed, h, x, n, y, J, H, (float4*)ws, dbg, peers));
    INF3DP_LAUNCH_CHECK("siren_rev_kernel");
    if (debug) {
        unsigned long long hbuf[256];
        INF3DP_CUDA(cudaMemcpyAsync(hbuf, dbg, sizeof(hbuf), cudaMemcpyDeviceToHost, stream));
        INF3DP_CUDA(cudaStreamSynchronize(stream));
        INF3DP_CUDA(cudaFree(dbg));
        const double u = hbuf[3] ? (double)hbuf[3] : 1.0;
        if (KIND == 0 && hbuf[64] != 0) {   // timeline of unit 4 of block 0 (cycles since its first MMA): MMA issue per K-step, epilogue warp 0
            const unsigned long long t0 = hbuf[64];
            for (int g = 0; g < 6; ++g) {
                fprintf(stderr, "[inf3dp rev trace] GEMM %d K-step issued at:", g);
                for (int t = 0; t < 16; ++t) fprintf(stderr, " %lld", (long long)(hbuf[64 + g * 16 + t] - t0));
                fprintf(stderr, "\n[inf3dp rev trace] GEMM %d accumulator seen by warp 0 at %lld, its slices stored at:", g, (long long)(hbuf[160 + g] - t0));
                for (int j = 0; j < 8; ++j) fprintf(stderr, " %lld", (long long)(hbuf[168 + g * 8 + j] - t0));
                fprintf(stderr, "\n");
            }
            fprintf(stderr, "[inf3dp rev trace] GEMM 2: first slice stored + published by warps 0..15 at:");
            for (int w = 0; w < 16; ++w) fprintf(stderr, " %lld", (long long)(hbuf[216 + w] - t0));
            fprintf(stderr, "\n");
        }
        for (int g = 0; g < 6; ++g)
            fprintf(stderr, "[inf3dp rev debug] GEMM %d: a_ready wait per round %.0f %.0f %.0f %.0f cyc\n", g, hbuf[32 + g * 4] / u,
                    hbuf[33 + g * 4] / u, hbuf[34 + g * 4] / u, hbuf[35 + g * 4] / u);
        fprintf(stderr, "[inf3dp rev debug] CTAS=%d KIND=%d block0: %.0f units, %.0f cyc/unit (%d GEMMs); MMA thread blocked on a_ready %.0f, "
                        "on w_full %.0f cyc/unit; epilogue warp0 blocked on d_full %.0f, warp5 %.0f, warp15 %.0f cyc/unit\n",
                CTAS, KIND, u, hbuf[0] / u, (KIND == 0 || KIND == 3) ? 6 : ((KIND == 1 || KIND == 4) ? 3 : 1), hbuf[1] / u, hbuf[2] / u, hbuf[8] / u, hbuf[13] / u, hbuf[23] / u);
    }
    return INF3DP_OK;
}

}  // namespace

size_t siren_rev_section_bytes() { return kRevSectionBytes; }

size_t siren_mlp_section_bytes() { return kMlpSectionBytes; }

int siren_mlp_pack(const float* const* W, const float* const* b, int out_features, void* packed, const SirenLayout& lay,
                   SirenHeader& hdr, cudaStream_t stream) {
    char* rv = (char*)packed + lay.off_rev;
    INF3DP_CUDA(cudaMemsetAsync(rv, 0, kMlpOffW2, stream));
    uint32_t* mx = reinterpret_cast<uint32_t*>(rv + kMlpOffMisc + 64);
    mlp_maxabs_kernel<<<64, 256, 0, stream>>>(W[1], 256 * 256, mx);
    INF3DP_LAUNCH_CHECK("mlp_maxabs_kernel");
    rev_pack_gemm_kernel<<<256, 256, 0, stream>>>(W[1], 1.f, mx, 0, rv + kMlpOffImg);
    INF3DP_LAUNCH_CHECK("rev_pack_gemm_kernel");
    mlp_pack_out_kernel<<<(256 * kMlpMaxOut + 255) / 256, 256, 0, stream>>>(W[2], b[2], out_features, rv);
    INF3DP_LAUNCH_CHECK("mlp_pack_out_kernel");
    hdr.mlp_ok = 1;
    return INF3DP_OK;
}

int siren_mlp_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, cudaStream_t stream) {
    INF3DP_CHECK_ARG(h.mlp_ok && h.off_rev != 0, "siren_mlp: packed weights carry no classifier operand image");
    INF3DP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "siren_mlp: x must be 16-byte aligned ([n,4] rows are read as float4)");
    return launch_rev<2, 2>(packed, h, x, n, y, nullptr, nullptr, nullptr, stream);
}

size_t siren_rev_workspace_bytes() { return (size_t)num_sms() * kStashBytesPerCta; }

// maxabs_bits: device uint32[4] = bit patterns of max |W| of layers 1..3 and of the output layer (filled by the forward
// engine's pack step, which runs first on the same stream).
int siren_rev_pack(const float* const* W, float omega0, const uint32_t* maxabs_bits, void* packed, const SirenLayout& lay,
                   SirenHeader& hdr, cudaStream_t stream) {
    char* rv = (char*)packed + lay.off_rev;
    for (int g = 0; g < 6; ++g) {
        const int layer = g < 3 ? 1 + g : 6 - g;   // forward: layers 1,2,3; backward: layers 3,2,1
        rev_pack_gemm_kernel<<<256, 256, 0, stream>>>(W[layer], omega0, maxabs_bits + (layer - 1), g >= 3 ? 1 : 0,
                                                      rv + kRevOffImg + (size_t)g * kGemmImgBytes);
        INF3DP_LAUNCH_CHECK("rev_pack_gemm_kernel");
    }
    for (int l = 1; l <= 3; ++l) {
        rev_colnorm_kernel<<<1, 256, 0, stream>>>(W[l], omega0, reinterpret_cast<float*>(rv + kRevOffMisc) + 8 + (l - 1));
        INF3DP_LAUNCH_CHECK("rev_colnorm_kernel");
    }
    rev_pack_out_kernel<<<1, 256, 0, stream>>>(W[4], maxabs_bits + 3, rv);
    INF3DP_LAUNCH_CHECK("rev_pack_out_kernel");
    hdr.rev_ok = 1;
    return INF3DP_OK;
}

int siren_rev_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, float* J, void* ws,
                     size_t ws_bytes, cudaStream_t stream) {
    INF3DP_CHECK_ARG(h.rev_ok && h.off_rev != 0, "siren_rev: packed weights carry no reverse-mode section");
    INF3DP_CHECK_ARG(ws != nullptr && ws_bytes >= siren_rev_workspace_bytes(),
                     "siren_rev: workspace of %zu bytes needed (inf3dp_siren_jet_workspace_bytes), got %zu",
                     siren_rev_workspace_bytes(), ws_bytes);
    static const int ctas = [] { const char* e = getenv("INF3DP_REV_CTAS"); return e && atoi(e) == 1 ? 1 : 2; }();
    return ctas == 1 ? launch_rev<1, 0>(packed, h, x, n, y, J, nullptr, ws, stream)
                     : launch_rev<2, 0>(packed, h, x, n, y, J, nullptr, ws, stream);
}

// Same query with the outputs packed as rows (value, d/dx, d/dy, d/dz) of a [n,4] tensor.
int siren_rev_launch_rows(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* rows, void* ws,
                          size_t ws_bytes, cudaStream_t stream) {
    INF3DP_CHECK_ARG(h.rev_ok && h.off_rev != 0, "siren_rev: packed weights carry no reverse-mode section");
    INF3DP_CHECK_ARG(ws != nullptr && ws_bytes >= siren_rev_workspace_bytes(),
                     "siren_rev: workspace of %zu bytes needed (inf3dp_siren_jet_workspace_bytes), got %zu",
                     siren_rev_workspace_bytes(), ws_bytes);
    INF3DP_CHECK_ARG((reinterpret_cast<uintptr_t>(rows) & 15) == 0, "siren_rev: rows must be 16-byte aligned");
    return launch_rev<2, 0>(packed, h, x, n, nullptr, nullptr, rows, ws, stream);
}

// Same rows, stored by the kernel's epilogue directly into the gathered buffers of the ranks of a multi-GPU job: the
// query fused with its all-gather (the one collective of the sharded path).
int siren_rev_launch_rows_peers(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* const* peer_rows,
                                int n_peers, float* multicast_rows, void* ws, size_t ws_bytes, cudaStream_t stream) {
    INF3DP_CHECK_ARG(h.rev_ok && h.off_rev != 0, "siren_rev: packed weights carry no reverse-mode section");
    INF3DP_CHECK_ARG(ws != nullptr && ws_bytes >= siren_rev_workspace_bytes(),
                     "siren_rev: workspace of %zu bytes needed (inf3dp_siren_jet_workspace_bytes), got %zu",
                     siren_rev_workspace_bytes(), ws_bytes);
    INF3DP_CHECK_ARG(multicast_rows != nullptr || (n_peers >= 1 && n_peers <= kMaxRowPeers && peer_rows != nullptr),
                     "siren_rev: need a multicast address or 1..%d peer row buffers", kMaxRowPeers);
    RowPeers peers{};
    peers.mc = reinterpret_cast<float4*>(multicast_rows);
    INF3DP_CHECK_ARG((reinterpret_cast<uintptr_t>(multicast_rows) & 15) == 0, "siren_rev: multicast rows must be 16-byte aligned");
    if (multicast_rows == nullptr) {
        peers.n = n_peers;
        for (int g = 0; g < n_peers; ++g) {
            INF3DP_CHECK_ARG(peer_rows[g] != nullptr && (reinterpret_cast<uintptr_t>(peer_rows[g]) & 15) == 0,
                             "siren_rev: peer row buffer %d is null or not 16-byte aligned", g);
            peers.dst[g] = reinterpret_cast<float4*>(peer_rows[g]);
        }
    }
    // a non-null Hout selects the rows epilogue; it is never dereferenced when peers are given
    float* tag = multicast_rows != nullptr ? multicast_rows : peer_rows[0];
    return launch_rev<2, 0>(packed, h, x, n, nullptr, nullptr, tag, ws, stream, peers);
}

// Value + gradient + Hessian (order 2) of single-output networks: forward-mode second-order jet on the CTA-pair engine.
int siren_jet2_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, float* J, float* H,
                      cudaStream_t stream) {
    INF3DP_CHECK_ARG(h.rev_ok && h.off_rev != 0, "siren_jet2: packed weights carry no CTA-pair operand images");
    static const int ctas = [] { const char* e = getenv("INF3DP_REV_CTAS"); return e && atoi(e) == 1 ? 1 : 2; }();
    return ctas == 1 ? launch_rev<1, 1>(packed, h, x, n, y, J, H, nullptr, stream)
                     : launch_rev<2, 1>(packed, h, x, n, y, J, H, nullptr, stream);
}

// Value only (order 0) of single-output field networks on the CTA-pair engine (KIND 4).
int siren_value_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, cudaStream_t stream) {
    INF3DP_CHECK_ARG(h.rev_ok && h.off_rev != 0, "siren_value: packed weights carry no CTA-pair operand images");
    return launch_rev<2, 4>(packed, h, x, n, y, nullptr, nullptr, nullptr, stream);
}

// The same outputs by forward-over-reverse differentiation (KIND 3): needs the reverse engine's scratch.
int siren_hess_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, float* J, float* H,
                      void* ws, size_t ws_bytes, cudaStream_t stream) {
    INF3DP_CHECK_ARG(h.rev_ok && h.off_rev != 0, "siren_hess: packed weights carry no reverse-mode section");
    INF3DP_CHECK_ARG(ws != nullptr && ws_bytes >= siren_rev_workspace_bytes(),
                     "siren_hess: workspace of %zu bytes needed (inf3dp_siren_jet_workspace_bytes), got %zu",
                     siren_rev_workspace_bytes(), ws_bytes);
    return launch_rev<2, 3>(packed, h, x, n, y, J, H, ws, stream);
}

}  // namespace inf3dp
