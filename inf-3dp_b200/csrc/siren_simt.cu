// FP32 CUDA-core engine of the fused SIREN jet (value / +Jacobian / +Hessian; sine or ReLU).
//
// One kernel per batch: a CTA owns a tile of P points and carries, for every point, C "jet columns"
// (C = 1: value; 4: value + d/dx,d/dy,d/dz; 10: + the 6 unique second derivatives) as C rows of a
// [128 x 256] FP32 activation matrix kept in shared memory.  A hidden layer is one 128x256x256 SGEMM
// (weights streamed from L2 through a double-buffered cp.async ring) followed by an in-place elementwise
// jet update -- activations never touch HBM and no autograd graph exists.
//
// Replaces modules.py:143-160 (forward) + diff_operators.py:128-132 (gradient) + :27-44 (hessian3d).
// Recurrences (omega0 folded into W', b' at pack time):
//   layer 0:  z = W0' x + b0';  a = sin z;  da_i = cos z * W0'[:,i];  d2a_ij = -sin z * W0'[:,i] W0'[:,j]
//   hidden:   z = W' a + b';  dz_i = W' da_i;  d2z_ij = W' d2a_ij
//             a = sin z;  da_i = cos z dz_i;  d2a_ij = cos z d2z_ij - sin z dz_i dz_j
//   output:   y = W a + b;  J_i = W da_i;  H_ij = W d2a_ij
// This engine is the accuracy reference for the tensor-core engine and serves every configuration the
// tensor-core engine does not (Hessian order, ReLU level MLP, in_features 4, out > 8, other depths).
#include "common.cuh"

namespace inf3dp {

namespace {

constexpr int kRows = 128;            // rows of the activation tile (points x jet columns)
constexpr int kAStride = 260;         // floats per activation row: 16B aligned, rows land on distinct bank groups
constexpr int kKChunk = 32;           // k-rows of W^T staged per pipeline stage
constexpr int kThreads = 256;
constexpr size_t kSmemA = (size_t)kRows * kAStride * sizeof(float);            // 133120
constexpr size_t kSmemW = (size_t)2 * kKChunk * kHidden * sizeof(float);       // 65536
constexpr size_t kSmemBytes = kSmemA + kSmemW;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Stage `rows` x 256 floats (contiguous in global memory) into shared memory with 16-byte cp.async.
__device__ __forceinline__ void stage_rows(float* dst, const float* src, int nfloat4) {
    for (int i = threadIdx.x; i < nfloat4; i += kThreads) cp_async16(dst + 4 * i, src + 4 * i);
}

template <int C, bool RELU>
__device__ __forceinline__ void jet_activation(float* As, int P, const float* __restrict__ bias) {
    // thread = feature; loops over the tile's points.  In-place on the pre-activation rows.
    const int n = threadIdx.x;
    const float bn = bias[n];
    for (int p = 0; p < P; ++p) {
        float* r = As + (size_t)(p * C) * kAStride + n;
        float z = r[0] + bn;
        if (RELU) {
            const bool on = z > 0.f;
            r[0] = on ? z : 0.f;
            if (C >= 4) {
#pragma unroll
                for (int c = 1; c < C; ++c) r[c * kAStride] = on ? r[c * kAStride] : 0.f;
            }
        } else {
            float s, c;
            sincosf(z, &s, &c);
            r[0] = s;
            if (C >= 4) {
                float d0 = r[1 * kAStride], d1 = r[2 * kAStride], d2 = r[3 * kAStride];
                r[1 * kAStride] = c * d0;
                r[2 * kAStride] = c * d1;
                r[3 * kAStride] = c * d2;
                if (C == 10) {
                    // order: 00 01 02 11 12 22
                    r[4 * kAStride] = c * r[4 * kAStride] - s * d0 * d0;
                    r[5 * kAStride] = c * r[5 * kAStride] - s * d0 * d1;
                    r[6 * kAStride] = c * r[6 * kAStride] - s * d0 * d2;
                    r[7 * kAStride] = c * r[7 * kAStride] - s * d1 * d1;
                    r[8 * kAStride] = c * r[8 * kAStride] - s * d1 * d2;
                    r[9 * kAStride] = c * r[9 * kAStride] - s * d2 * d2;
                }
            }
        }
    }
}

template <int C, bool RELU>
__global__ void __launch_bounds__(kThreads, 1)
siren_simt_kernel(const char* __restrict__ packed, SirenHeader h, const float* __restrict__ x, int64_t n,
                  float* __restrict__ y, float* __restrict__ J, float* __restrict__ H) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);
    float* Ws = reinterpret_cast<float*>(smem_raw + kSmemA);

    constexpr int P = kRows / C;  // points per tile (128, 32, 12)
    const float* w0 = reinterpret_cast<const float*>(packed + h.off_w0);
    const float* b0 = reinterpret_cast<const float*>(packed + h.off_b0);
    const float* wt = reinterpret_cast<const float*>(packed + h.off_wt);
    const float* bh = reinterpret_cast<const float*>(packed + h.off_b);
    const float* wout = reinterpret_cast<const float*>(packed + h.off_wout);
    const float* bout = reinterpret_cast<const float*>(packed + h.off_bout);
    const int L = h.num_hidden_layers;
    const int in_f = h.in_features;
    const int out_f = h.out_features;

    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int64_t num_tiles = (n + P - 1) / P;

    // rows P*C .. 127 are never read back as results; zero them once so the GEMM reads finite values
    for (int i = tid; i < (kRows - P * C) * kAStride; i += kThreads) As[(size_t)P * C * kAStride + i] = 0.f;

    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t base = tile * P;
        const int valid = (int)min((int64_t)P, n - base);
        __syncthreads();  // previous tile's output phase is done with As/Ws

        // ---- layer 0 (FP32, no GEMM): thread = feature ------------------------------------------------
        {
            const float4 w = reinterpret_cast<const float4*>(w0)[tid];
            const float bb = b0[tid];
            for (int p = 0; p < P; ++p) {
                float* r = As + (size_t)(p * C) * kAStride + tid;
                float z = bb;
                if (p < valid) {
                    const float* xp = x + (base + p) * in_f;
                    z = fmaf(w.x, xp[0], z);
                    z = fmaf(w.y, xp[1], z);
                    z = fmaf(w.z, xp[2], z);
                    if (in_f == 4) z = fmaf(w.w, xp[3], z);
                }
                if (RELU) {
                    const bool on = z > 0.f;
                    r[0] = on ? z : 0.f;
                    if (C >= 4) {
                        r[1 * kAStride] = on ? w.x : 0.f;
                        r[2 * kAStride] = on ? w.y : 0.f;
                        r[3 * kAStride] = on ? w.z : 0.f;
                    }
                    if (C == 10) {
#pragma unroll
                        for (int q = 4; q < 10; ++q) r[q * kAStride] = 0.f;
                    }
                } else {
                    float s, c;
                    sincosf(z, &s, &c);
                    r[0] = s;
                    if (C >= 4) {
                        r[1 * kAStride] = c * w.x;
                        r[2 * kAStride] = c * w.y;
                        r[3 * kAStride] = c * w.z;
                    }
                    if (C == 10) {
                        r[4 * kAStride] = -s * w.x * w.x;
                        r[5 * kAStride] = -s * w.x * w.y;
                        r[6 * kAStride] = -s * w.x * w.z;
                        r[7 * kAStride] = -s * w.y * w.y;
                        r[8 * kAStride] = -s * w.y * w.z;
                        r[9 * kAStride] = -s * w.z * w.z;
                    }
                }
            }
        }
        __syncthreads();

        // ---- hidden layers: Z[128,256] = A[128,256] * Wt[256,256] ---------------------------------------
        for (int l = 0; l < L; ++l) {
            const float* wl = wt + (size_t)l * kHidden * kHidden;
            float acc[8][16];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;

            constexpr int kChunks = kHidden / kKChunk;  // 8
            stage_rows(Ws, wl, kKChunk * kHidden / 4);
            cp_async_commit();
            for (int ch = 0; ch < kChunks; ++ch) {
                if (ch + 1 < kChunks) {
                    stage_rows(Ws + ((ch + 1) & 1) * kKChunk * kHidden, wl + (size_t)(ch + 1) * kKChunk * kHidden,
                               kKChunk * kHidden / 4);
                    cp_async_commit();
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                const float* wb = Ws + (ch & 1) * kKChunk * kHidden;
                const float* ab = As + (size_t)(ty * 8) * kAStride + ch * kKChunk;
#pragma unroll 2
                for (int kk = 0; kk < kKChunk; kk += 4) {
                    float4 a[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(ab + (size_t)i * kAStride + kk);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 w[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            w[j] = *reinterpret_cast<const float4*>(wb + (kk + q) * kHidden + 64 * j + 4 * tx);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float av = q == 0 ? a[i].x : q == 1 ? a[i].y : q == 2 ? a[i].z : a[i].w;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc[i][4 * j + 0] = fmaf(av, w[j].x, acc[i][4 * j + 0]);
                                acc[i][4 * j + 1] = fmaf(av, w[j].y, acc[i][4 * j + 1]);
                                acc[i][4 * j + 2] = fmaf(av, w[j].z, acc[i][4 * j + 2]);
                                acc[i][4 * j + 3] = fmaf(av, w[j].w, acc[i][4 * j + 3]);
                            }
                        }
                    }
                }
                __syncthreads();  // everyone is done with this W stage (and, on the last chunk, with As)
            }
            // write pre-activations over the (now dead) inputs
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<float4*>(As + (size_t)(ty * 8 + i) * kAStride + 64 * j + 4 * tx) =
                        make_float4(acc[i][4 * j], acc[i][4 * j + 1], acc[i][4 * j + 2], acc[i][4 * j + 3]);
            __syncthreads();
            jet_activation<C, RELU>(As, P, bh + (size_t)l * kHidden);
            __syncthreads();
        }

        // ---- output layer (linear): one dot product per (row, output) -----------------------------------
        {
            const int nstage = out_f * kHidden / 4;
            for (int i = tid; i < nstage; i += kThreads)
                reinterpret_cast<float4*>(Ws)[i] = reinterpret_cast<const float4*>(wout)[i];
            __syncthreads();
            const int total = kRows * out_f;
            for (int idx = tid; idx < total; idx += kThreads) {
                const int r = idx & (kRows - 1);
                const int o = idx >> 7;
                const int p = r / C, c = r - p * C;
                if (p >= valid) continue;
                const float* ar = As + (size_t)r * kAStride;
                const float* wr = Ws + o * kHidden;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
                for (int k = 0; k < kHidden; k += 4) {
                    const float4 a = *reinterpret_cast<const float4*>(ar + k);
                    const float4 w = *reinterpret_cast<const float4*>(wr + k);
                    s0 = fmaf(a.x, w.x, s0);
                    s1 = fmaf(a.y, w.y, s1);
                    s2 = fmaf(a.z, w.z, s2);
                    s3 = fmaf(a.w, w.w, s3);
                }
                const float v = (s0 + s1) + (s2 + s3);
                const int64_t po = (base + p) * out_f + o;
                if (c == 0) {
                    y[po] = v + bout[o];
                } else if (c < 4) {
                    J[po * 3 + (c - 1)] = v;
                } else {
                    // 00 01 02 11 12 22 -> symmetric 3x3
                    float* hp = H + po * 9;
                    switch (c) {
                        case 4: hp[0] = v; break;
                        case 5: hp[1] = v; hp[3] = v; break;
                        case 6: hp[2] = v; hp[6] = v; break;
                        case 7: hp[4] = v; break;
                        case 8: hp[5] = v; hp[7] = v; break;
                        default: hp[8] = v; break;
                    }
                }
            }
        }
    }
}

template <int C, bool RELU>
int launch_one(const void* packed, const SirenHeader& h, const float* x, int64_t n, float* y, float* J, float* H,
               cudaStream_t stream) {
    auto kern = siren_simt_kernel<C, RELU>;
    // per device and idempotent; cheap enough to repeat on every launch
    INF3DP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    constexpr int P = kRows / C;
    const int64_t tiles = (n + P - 1) / P;
    const int grid = (int)(tiles < (int64_t)siren_sms() ? tiles : (int64_t)siren_sms());
    kern<<<grid, kThreads, kSmemBytes, stream>>>((const char*)packed, h, x, n, y, J, H);
    INF3DP_LAUNCH_CHECK("siren_simt_kernel");
    return INF3DP_OK;
}

}  // namespace

int siren_simt_launch(const void* packed, const SirenHeader& h, const float* x, int64_t n, int order, float* y,
                      float* J, float* H, cudaStream_t stream) {
    const bool relu = h.activation == INF3DP_ACT_RELU;
    if (order == 0) return relu ? launch_one<1, true>(packed, h, x, n, y, J, H, stream)
                                : launch_one<1, false>(packed, h, x, n, y, J, H, stream);
    if (order == 1) return relu ? launch_one<4, true>(packed, h, x, n, y, J, H, stream)
                                : launch_one<4, false>(packed, h, x, n, y, J, H, stream);
    return relu ? launch_one<10, true>(packed, h, x, n, y, J, H, stream)
                : launch_one<10, false>(packed, h, x, n, y, J, H, stream);
}

}  // namespace inf3dp
