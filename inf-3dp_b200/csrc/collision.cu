// Fused predicate / compaction stages of the TV-SDF collision query.
//
// Reference: level_collision.CollTVSDF.forward (level_collision.py:73-237) on the nozzle clouds posed by
// utils.transform_nozzle_pcd_with_quaternion (utils.py:612-621).  The reference -- and round 1 of this build -- glue the
// field sweeps together with boolean masks, nonzero / gather / scatter, repeat_interleave and torch.where in eager PyTorch
// (dozens of elementwise launches and index tensors over N*M points).  Here every step between two field sweeps is ONE
// kernel that evaluates the predicate of that step and compacts the surviving points with warp-aggregated appends
// (one atomic per warp), and a point that is finished -- outside the cube, outside the part, not colliding, or resolved --
// reports its result at that moment (per point for the CollTVSDF surface, or reduced per waypoint with atomicMax /
// flag stores for the sweep: the [N,M] arrays are then never materialised, nor is the posed cloud [N,M,3]).
//
//   pose      waypoint + quaternion + nozzle point -> posed point; cube test (:92); base-plate rule (:113-117)   -> cube list
//   [sweep]   SDF value on the cube list (the gradient is needed for colliding points only: queried lazily)
//   inside    sdf < 0 (:120)                                                                                    -> inside list
//   [sweep]   slice value + gradient on the inside list;  pack (xyz, slice) -> level MLP logits
//   predicate level arg-max, collision predicate (:143-148), Eq. 27 (:169-172), recheck split (:173),
//             first push along the slice gradient (:202-203)                                                    -> recheck list
//   [sweep]   slice gradient at the pushed points
//   push      second push (:206-208)
//   [sweep]   slice value at the twice-pushed points;  pack -> level MLP logits
//   alter     alter predicate (:216-219), final TV-SDF value and which gradient it carries
//   [sweep]   (CollTVSDF surface only) SDF value + gradient on the points whose result carries the SDF gradient; scatter
//
// HBM-bound elementwise work: coalesced reads of the compact lists, 16-byte row loads, one pass per stage.
#include "common.cuh"

namespace inf3dp {
namespace {

constexpr int kBaseFlag = 0x40000000;   // bit 30 of a point id: the point lies under the base plate (z < base_th)
constexpr int kIdMask = 0x3fffffff;
constexpr int kThreads = 256;

struct CollOut {
    float* depth_p;       // [P]   per point (CollTVSDF surface), pre-zeroed; or null
    uint8_t* mask_p;      // [P]
    float* grads_p;       // [P,3]
    unsigned* depth_w;    // [Nw]  per waypoint (sweep), pre-zeroed: max depth as float bits (depths are >= 0); or null
    uint8_t* hit_w;       // [Nw]
    int M;
};

// A finished point: depth >= 0 (or NaN), in collision.  grad: the direction its depth carries, or null when it is filled in
// later (lazy SDF gradient).
__device__ __forceinline__ void report(const CollOut& o, int id, float depth, const float* grad) {
    if (o.depth_p != nullptr) {
        o.depth_p[id] = depth;
        o.mask_p[id] = 1;
        if (grad != nullptr) { o.grads_p[3ll * id] = grad[0]; o.grads_p[3ll * id + 1] = grad[1]; o.grads_p[3ll * id + 2] = grad[2]; }
    }
    if (o.depth_w != nullptr) {
        const int w = id / o.M;
        atomicMax(&o.depth_w[w], __float_as_uint(depth > 0.f ? depth : 0.f));   // a NaN depth counts as 0 here, like max()
        o.hit_w[w] = 1;
    }
}

__device__ __forceinline__ void report_base(const CollOut& o, int id) {   // level_collision.py:113-117: depth -(-0.1), z-up
    const float up[3] = {0.f, 0.f, 1.f};
    report(o, id, 0.1f, up);
}

// Warp-aggregated append: every lane of the warp calls it; returns the slot of a lane with pred, -1 otherwise.
__device__ __forceinline__ int warp_append(bool pred, int* counter) {
    const unsigned bal = __ballot_sync(0xffffffffu, pred);
    if (bal == 0) return -1;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(bal) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pred ? base + __popc(bal & ((1u << lane) - 1u)) : -1;
}

__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }

// F.normalize(g, dim=-1): g / max(||g||, 1e-12)
__device__ __forceinline__ void normalize3(const float* g, float* out) {
    const float nrm = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    const float d = fmaxf(nrm, 1e-12f);
    out[0] = g[0] / d; out[1] = g[1] / d; out[2] = g[2] / d;
}

__device__ __forceinline__ int argmax_first(const float* __restrict__ logits, int C) {   // torch.argmax: first maximum
    int best = 0;
    float bv = logits[0];
    for (int c = 1; c < C; ++c) { const float v = logits[c]; if (v > bv || (v != v && bv == bv)) { bv = v; best = c; } }
    return best;
}

// ---- pose + cube ---------------------------------------------------------------------------------------------------
// quats != null: point (w, m) = rotate(normalize(quats[w]), nozzle[m]) + waypts[w]   (utils.py:587-594, :612-621)
// quats == null: `nozzle` is the posed cloud [Nw*M,3] itself (CollTVSDF surface)
__global__ void __launch_bounds__(kThreads)
coll_pose_kernel(const float* __restrict__ waypts, const float* __restrict__ quats, const float* __restrict__ nozzle,
                 long long P, int M, float base_th, float* __restrict__ cube_x, int* __restrict__ cube_id,
                 int* __restrict__ counters, CollOut o) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long rounds = (P + stride - 1) / stride;
    for (long long r = 0; r < rounds; ++r) {
        const long long p = r * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        bool in_cube = false, base = false;
        float x[3] = {0.f, 0.f, 0.f};
        if (p < P) {
            if (quats != nullptr) {
                const int w = (int)(p / M), m = (int)(p - (long long)w * M);
                const float4 q4 = *reinterpret_cast<const float4*>(quats + 4ll * w);
                const float qn = fmaxf(sqrtf(q4.x * q4.x + q4.y * q4.y + q4.z * q4.z + q4.w * q4.w), 1e-12f);
                const float qs = q4.x / qn, qx = q4.y / qn, qy = q4.z / qn, qz = q4.w / qn;
                const float vx = nozzle[3 * m], vy = nozzle[3 * m + 1], vz = nozzle[3 * m + 2];
                const float tx = 2.f * (qy * vz - qz * vy), ty = 2.f * (qz * vx - qx * vz), tz = 2.f * (qx * vy - qy * vx);
                x[0] = (vx + qs * tx + (qy * tz - qz * ty)) + waypts[3ll * w];
                x[1] = (vy + qs * ty + (qz * tx - qx * tz)) + waypts[3ll * w + 1];
                x[2] = (vz + qs * tz + (qx * ty - qy * tx)) + waypts[3ll * w + 2];
            } else {
                x[0] = nozzle[3 * p]; x[1] = nozzle[3 * p + 1]; x[2] = nozzle[3 * p + 2];
            }
            in_cube = x[0] >= -1.f && x[0] <= 1.f && x[1] >= -1.f && x[1] <= 1.f && x[2] >= -1.f && x[2] <= 1.f;   // :92
            base = x[2] < base_th;                                                                                  // :113
        }
        const int slot = warp_append(in_cube, &counters[0]);
        if (slot >= 0) {
            cube_x[3ll * slot] = x[0]; cube_x[3ll * slot + 1] = x[1]; cube_x[3ll * slot + 2] = x[2];
            cube_id[slot] = (int)p | (base ? kBaseFlag : 0);
        } else if (base) {
            report_base(o, (int)p);
        }
    }
}

// ---- inside filter ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
coll_inside_kernel(const float* __restrict__ sdf, long long n, const float* __restrict__ cube_x, const int* __restrict__ cube_id,
                   float* __restrict__ in_x, int* __restrict__ in_id, float* __restrict__ in_sdf, int* __restrict__ counters,
                   CollOut o) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long rounds = (n + stride - 1) / stride;
    for (long long r = 0; r < rounds; ++r) {
        const long long k = r * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        bool inside = false;
        int id = 0;
        float v = 0.f;
        if (k < n) { v = sdf[k]; id = cube_id[k]; inside = v < 0.f; }                                                // :120
        const int slot = warp_append(inside, &counters[1]);
        if (slot >= 0) {
            in_x[3ll * slot] = cube_x[3 * k]; in_x[3ll * slot + 1] = cube_x[3 * k + 1]; in_x[3ll * slot + 2] = cube_x[3 * k + 2];
            in_id[slot] = id;
            in_sdf[slot] = v;
        } else if (k < n && (id & kBaseFlag)) {
            report_base(o, id & kIdMask);
        }
    }
}

// (x, y, z, v) rows for the level MLP (level_collision.py:139-140, :212): v = values[k * vstride]
__global__ void __launch_bounds__(kThreads)
coll_pack4_kernel(const float* __restrict__ x3, const float* __restrict__ values, int vstride, long long n, float4* __restrict__ x4) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) x4[k] = make_float4(x3[3 * k], x3[3 * k + 1], x3[3 * k + 2], values[k * vstride]);
}

// ---- collision predicate, Eq. 27, recheck split, first push -----------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
coll_predicate_kernel(const float4* __restrict__ rows, const float* __restrict__ logits, int C, long long n,
                      const float* __restrict__ in_x, const int* __restrict__ in_id, const float* __restrict__ in_sdf,
                      const int* __restrict__ levels_w, const float* __restrict__ maxsl, float* __restrict__ re_x,
                      int* __restrict__ re_src, float* __restrict__ re_dist, int* __restrict__ grad_src, int want_grads,
                      int* __restrict__ counters, CollOut o) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long rounds = (n + stride - 1) / stride;
    for (long long r = 0; r < rounds; ++r) {
        const long long k = r * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        bool recheck = false, need_sdf_grad = false;
        float push[3] = {0.f, 0.f, 0.f}, dist = 0.f;
        if (k < n) {
            const int idf = in_id[k], id = idf & kIdMask;
            const int w = id / o.M;
            const float4 row = rows[k];                       // slice value, slice gradient
            const int level = argmax_first(logits + k * C, C);
            const int wl = levels_w[w];
            const float wmax = maxsl[(long long)w * C + wl];
            const bool coll = (level < wl) || (level == wl && row.x < wmax);                                      // :143-146
            if (!coll) {
                if (idf & kBaseFlag) report_base(o, id);
            } else {
                const float sdfv = in_sdf[k];
                const float g[3] = {row.y, row.z, row.w};
                const float nrm = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
                dist = (row.x - maxsl[(long long)w * C + level]) / nrm;                                            // :166-169 (Eq. 27)
                recheck = dist > sdfv;                                                                             // :173
                if (recheck) {
                    float gn[3];
                    normalize3(g, gn);
                    const float a = fabsf(dist);
#pragma unroll
                    for (int d = 0; d < 3; ++d) push[d] = clamp1(in_x[3 * k + d] + a * gn[d]);                      // :202-203
                } else {
                    const float tv = (dist != dist) ? dist : fmaxf(sdfv, dist);                                    // torch.max keeps NaN
                    report(o, id, fmaxf(-tv, 0.f) + (tv != tv ? tv : 0.f), nullptr);                               // relu(-tvsdf)
                    need_sdf_grad = want_grads != 0;
                }
            }
        }
        const int slot = warp_append(recheck, &counters[2]);
        if (slot >= 0) {
            re_x[3ll * slot] = push[0]; re_x[3ll * slot + 1] = push[1]; re_x[3ll * slot + 2] = push[2];
            re_src[slot] = (int)k;
            re_dist[slot] = dist;
        }
        const int gslot = warp_append(need_sdf_grad, &counters[3]);
        if (gslot >= 0) grad_src[gslot] = (int)k;
    }
}

// ---- second push -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
coll_push_kernel(const float4* __restrict__ rows2, long long n, const float* __restrict__ re_x, const float* __restrict__ re_dist,
                 float* __restrict__ push_x) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 row = rows2[k];
    const float g[3] = {row.y, row.z, row.w};
    float gn[3];
    normalize3(g, gn);
    const float step = fmaxf(fabsf(re_dist[k]) * 0.5f, 0.04f);                                                      // :206
#pragma unroll
    for (int d = 0; d < 3; ++d) push_x[3 * k + d] = clamp1(re_x[3 * k + d] + step * gn[d]);                          // :207-208
}

// ---- alter predicate ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
coll_alter_kernel(const float* __restrict__ y2, const float* __restrict__ logits2, int C, long long n,
                  const int* __restrict__ re_src, const float* __restrict__ re_dist, const float4* __restrict__ rows,
                  const int* __restrict__ in_id, const float* __restrict__ in_sdf, const int* __restrict__ levels_w,
                  const float* __restrict__ maxsl, int* __restrict__ grad_src, int want_grads, int* __restrict__ counters,
                  CollOut o) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long rounds = (n + stride - 1) / stride;
    for (long long r = 0; r < rounds; ++r) {
        const long long k = r * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        bool need_sdf_grad = false;
        int src = 0;
        if (k < n) {
            src = re_src[k];
            const int id = in_id[src] & kIdMask;
            const int w = id / o.M;
            const int wl = levels_w[w];
            const float wmax = maxsl[(long long)w * C + wl];
            const int level2 = argmax_first(logits2 + k * C, C);
            const bool alter = (level2 < wl) || (y2[k] < wmax && level2 == wl);                                    // :216-217
            const float fin = alter ? in_sdf[src] : re_dist[k];                                                    // :218
            if (alter) {
                report(o, id, fmaxf(-fin, 0.f), nullptr);
                need_sdf_grad = want_grads != 0;
            } else {
                const float4 row = rows[src];
                const float g[3] = {row.y, row.z, row.w};                                                          // :219 slice gradient
                report(o, id, fmaxf(-fin, 0.f), g);
            }
        }
        const int gslot = warp_append(need_sdf_grad, &counters[3]);
        if (gslot >= 0) grad_src[gslot] = src;
    }
}

// coordinates of the points whose result carries the SDF gradient (for the lazy SDF value + gradient sweep)
__global__ void __launch_bounds__(kThreads)
coll_gather_x_kernel(const int* __restrict__ grad_src, long long n, const float* __restrict__ in_x, float* __restrict__ gx) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int s = grad_src[k];
    gx[3 * k] = in_x[3ll * s]; gx[3 * k + 1] = in_x[3ll * s + 1]; gx[3 * k + 2] = in_x[3ll * s + 2];
}

__global__ void __launch_bounds__(kThreads)
coll_scatter_grads_kernel(const float4* __restrict__ rows3, long long n, const int* __restrict__ grad_src,
                          const int* __restrict__ in_id, float* __restrict__ grads_p) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int id = in_id[grad_src[k]] & kIdMask;
    const float4 row = rows3[k];
    grads_p[3ll * id] = row.y; grads_p[3ll * id + 1] = row.z; grads_p[3ll * id + 2] = row.w;
}

inline unsigned grid_for(long long n) {
    long long b = (n + kThreads - 1) / kThreads;
    const long long cap = (long long)num_sms() * 32;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}
inline unsigned blocks_for(long long n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace
}  // namespace inf3dp

using namespace inf3dp;

extern "C" {

#define COLL_OUT CollOut{depth_p, mask_p, grads_p, depth_w, hit_w, M}

int inf3dp_coll_pose(const float* waypts, const float* quats, const float* nozzle, int64_t Nw, int M, float base_th,
                     float* cube_x, int32_t* cube_id, int32_t* counters, float* depth_p, uint8_t* mask_p, float* grads_p,
                     uint32_t* depth_w, uint8_t* hit_w, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(Nw >= 0 && M >= 1, "coll_pose: bad sizes");
    const long long P = (long long)Nw * M;
    INF3DP_CHECK_ARG(P < (1ll << 30), "coll_pose: %lld points per call exceed 2^30 (chunk the waypoints)", P);
    INF3DP_CHECK_ARG(nozzle && cube_x && cube_id && counters, "coll_pose: null pointer");
    INF3DP_CHECK_ARG(quats == nullptr || (waypts != nullptr && (reinterpret_cast<uintptr_t>(quats) & 15) == 0),
                     "coll_pose: quaternions need waypoints and 16-byte alignment");
    INF3DP_CHECK_ARG((depth_p && mask_p && grads_p) || (depth_w && hit_w), "coll_pose: no output surface");
    cudaStream_t st = (cudaStream_t)stream;
    INF3DP_CUDA(cudaMemsetAsync(counters, 0, 8 * sizeof(int32_t), st));
    if (P == 0) return INF3DP_OK;
    coll_pose_kernel<<<grid_for(P), kThreads, 0, st>>>(waypts, quats, nozzle, P, M, base_th, cube_x, cube_id, counters, COLL_OUT);
    INF3DP_LAUNCH_CHECK("coll_pose_kernel");
    return INF3DP_OK;
}

int inf3dp_coll_inside(const float* sdf, int64_t n, const float* cube_x, const int32_t* cube_id, float* in_x, int32_t* in_id,
                       float* in_sdf, int32_t* counters, int M, float* depth_p, uint8_t* mask_p, float* grads_p,
                       uint32_t* depth_w, uint8_t* hit_w, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(n >= 0 && M >= 1, "coll_inside: bad sizes");
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(sdf && cube_x && cube_id && in_x && in_id && in_sdf && counters, "coll_inside: null pointer");
    coll_inside_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(sdf, n, cube_x, cube_id, in_x, in_id, in_sdf, counters, COLL_OUT);
    INF3DP_LAUNCH_CHECK("coll_inside_kernel");
    return INF3DP_OK;
}

int inf3dp_coll_pack4(const float* x3, const float* values, int value_stride, int64_t n, float* x4, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(n >= 0 && value_stride >= 1, "coll_pack4: bad sizes");
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(x3 && values && x4 && (reinterpret_cast<uintptr_t>(x4) & 15) == 0, "coll_pack4: null or unaligned pointer");
    coll_pack4_kernel<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(x3, values, value_stride, n, reinterpret_cast<float4*>(x4));
    INF3DP_LAUNCH_CHECK("coll_pack4_kernel");
    return INF3DP_OK;
}

int inf3dp_coll_predicate(const float* rows, const float* logits, int C, int64_t n, const float* in_x, const int32_t* in_id,
                          const float* in_sdf, const int32_t* levels_w, const float* maxsl, int M, float* re_x,
                          int32_t* re_src, float* re_dist, int32_t* grad_src, int want_grads, int32_t* counters,
                          float* depth_p, uint8_t* mask_p, float* grads_p, uint32_t* depth_w, uint8_t* hit_w,
                          inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(n >= 0 && C >= 1 && M >= 1, "coll_predicate: bad sizes");
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(rows && logits && in_x && in_id && in_sdf && levels_w && maxsl && re_x && re_src && re_dist && counters,
                     "coll_predicate: null pointer");
    INF3DP_CHECK_ARG(!want_grads || grad_src, "coll_predicate: want_grads needs grad_src");
    INF3DP_CHECK_ARG((reinterpret_cast<uintptr_t>(rows) & 15) == 0, "coll_predicate: rows must be 16-byte aligned");
    coll_predicate_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(rows), logits, C, n, in_x, in_id, in_sdf, levels_w, maxsl, re_x, re_src, re_dist, grad_src,
        want_grads, counters, COLL_OUT);
    INF3DP_LAUNCH_CHECK("coll_predicate_kernel");
    return INF3DP_OK;
}

int inf3dp_coll_push(const float* rows2, int64_t n, const float* re_x, const float* re_dist, float* push_x, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(n >= 0, "coll_push: bad sizes");
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(rows2 && re_x && re_dist && push_x && (reinterpret_cast<uintptr_t>(rows2) & 15) == 0, "coll_push: null or unaligned pointer");
    coll_push_kernel<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(rows2), n, re_x, re_dist, push_x);
    INF3DP_LAUNCH_CHECK("coll_push_kernel");
    return INF3DP_OK;
}

int inf3dp_coll_alter(const float* y2, const float* logits2, int C, int64_t n, const int32_t* re_src, const float* re_dist,
                      const float* rows, const int32_t* in_id, const float* in_sdf, const int32_t* levels_w, const float* maxsl,
                      int M, int32_t* grad_src, int want_grads, int32_t* counters, float* depth_p, uint8_t* mask_p,
                      float* grads_p, uint32_t* depth_w, uint8_t* hit_w, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(n >= 0 && C >= 1 && M >= 1, "coll_alter: bad sizes");
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(y2 && logits2 && re_src && re_dist && rows && in_id && in_sdf && levels_w && maxsl && counters, "coll_alter: null pointer");
    INF3DP_CHECK_ARG(!want_grads || grad_src, "coll_alter: want_grads needs grad_src");
    coll_alter_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(y2, logits2, C, n, re_src, re_dist,
                                                                           reinterpret_cast<const float4*>(rows), in_id, in_sdf, levels_w,
                                                                           maxsl, grad_src, want_grads, counters, COLL_OUT);
    INF3DP_LAUNCH_CHECK("coll_alter_kernel");
    return INF3DP_OK;
}

int inf3dp_coll_gather_x(const int32_t* grad_src, int64_t n, const float* in_x, float* gx, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(n >= 0, "coll_gather_x: bad sizes");
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(grad_src && in_x && gx, "coll_gather_x: null pointer");
    coll_gather_x_kernel<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(grad_src, n, in_x, gx);
    INF3DP_LAUNCH_CHECK("coll_gather_x_kernel");
    return INF3DP_OK;
}

int inf3dp_coll_scatter_grads(const float* rows3, int64_t n, const int32_t* grad_src, const int32_t* in_id, float* grads_p,
                              inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(n >= 0, "coll_scatter_grads: bad sizes");
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(rows3 && grad_src && in_id && grads_p && (reinterpret_cast<uintptr_t>(rows3) & 15) == 0, "coll_scatter_grads: null or unaligned pointer");
    coll_scatter_grads_kernel<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(rows3), n, grad_src, in_id, grads_p);
    INF3DP_LAUNCH_CHECK("coll_scatter_grads_kernel");
    return INF3DP_OK;
}

}  // extern "C"
