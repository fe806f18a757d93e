// Three-field voxel intersection for sm_100a.
//
// Replaces contour_cuda/ext.cpp:20-46 -> intersection.cu:227-287 (kernel :31-224) of the reference: for every
// voxel of a corner-expanded [Dx,Dy,Dz,8] grid and every (slice, lattice1, lattice2) isovalue triple whose
// values lie inside the voxel's [min,max] of the respective field, the cell (s,l1,l2) receives the mean over the
// three fields of the mean edge-crossing point on the 12 voxel edges.
//
// Differences in HOW (results are a valid outcome of the reference):
//   * the three isovalue arrays are rank-sorted once, so a voxel binary-searches its three index windows instead
//     of the reference's 128 + 128*h1 + Ns*h1*h2 linear scans
//   * the reference lets voxels that hit the same cell race (last writer wins, :119-124, :216-218); here pass 1
//     elects the LOWEST linear voxel index per cell with atomicMin and pass 2 -- run only over the compacted
//     list of voxels that hit anything -- lets the winner write.  Deterministic, one full read of the fields.
//   * outputs are initialised by the same call (mask 0, indices -1, points 0), 16-byte vectorised
#include "common.cuh"

namespace inf3dp {
namespace {

struct IsectWs {
    int* hdr;         // [0] number of voxels in the work list, [4..6] finite isovalue counts (slice, l1, l2)
    float* iso_sorted[3];
    int* iso_perm[3];
    int* winner;      // [cells]
    int* worklist;    // [voxels]
    size_t total_bytes;
};

IsectWs carve_isect(void* base, long long voxels, long long cells, int Ns, int N1, int N2) {
    IsectWs w{};
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    char* b = (char*)base;
    const int n[3] = {Ns, N1, N2};
    w.hdr = (int*)(b + take(64));
    for (int i = 0; i < 3; ++i) {
        w.iso_sorted[i] = (float*)(b + take((size_t)(n[i] > 0 ? n[i] : 1) * 4));
        w.iso_perm[i] = (int*)(b + take((size_t)(n[i] > 0 ? n[i] : 1) * 4));
    }
    w.winner = (int*)(b + take((size_t)cells * 4));
    w.worklist = (int*)(b + take((size_t)voxels * 4));
    w.total_bytes = o;
    return w;
}

// Rank sort (ties keep input order).  NaN isovalues are parked behind the finite ones and never searched:
// a NaN isovalue is ignored here (the reference would let it hit every voxel, intersection.cu:98 -- a
// documented divergence on meaningless input).
__device__ __forceinline__ void rank_sort(const float* __restrict__ iso, int n, float* sorted, int* perm, int* n_finite) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float vi = iso[i];
        int rank = 0;
        if (vi == vi) {
            for (int j = 0; j < n; ++j) { const float vj = iso[j]; rank += (vj < vi) || (vj == vi && j < i); }
        } else {
            int nn = 0, before = 0;
            for (int j = 0; j < n; ++j) { const float vj = iso[j]; nn += (vj == vj); before += (vj != vj) && j < i; }
            rank = nn + before;
        }
        sorted[rank] = vi;
        perm[rank] = i;
    }
    if (threadIdx.x == 0) {
        int nn = 0;
        for (int j = 0; j < n; ++j) nn += (iso[j] == iso[j]);
        *n_finite = nn;
    }
}

__global__ void isect_prep_kernel(const float* s_iso, const float* a_iso, const float* b_iso, int Ns, int N1, int N2,
                                  IsectWs w) {
    if (threadIdx.x < 4) w.hdr[threadIdx.x] = 0;
    rank_sort(s_iso, Ns, w.iso_sorted[0], w.iso_perm[0], &w.hdr[4]);
    rank_sort(a_iso, N1, w.iso_sorted[1], w.iso_perm[1], &w.hdr[5]);
    rank_sort(b_iso, N2, w.iso_sorted[2], w.iso_perm[2], &w.hdr[6]);
}

// Output + winner initialisation, flat over cells (4 cells per thread so every store is 4..16 bytes wide).
__global__ void isect_init_kernel(long long cells, int* __restrict__ winner, uint8_t* __restrict__ mask,
                                  int16_t* __restrict__ idx, float* __restrict__ pts) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // quad of cells
    const long long c0 = q * 4;
    if (c0 >= cells) return;
    if (c0 + 4 <= cells) {
        // torch allocations are >= 256-byte aligned, so these vector stores are aligned for any quad index
        *reinterpret_cast<int4*>(winner + c0) = make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff);
        *reinterpret_cast<uint32_t*>(mask + c0) = 0u;
        uint2* ip = reinterpret_cast<uint2*>(idx + 3 * c0);  // 12 int16 = 24 bytes
        ip[0] = make_uint2(0xffffffffu, 0xffffffffu);
        ip[1] = make_uint2(0xffffffffu, 0xffffffffu);
        ip[2] = make_uint2(0xffffffffu, 0xffffffffu);
        float4* pp = reinterpret_cast<float4*>(pts + 3 * c0);  // 12 floats = 48 bytes
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        pp[0] = z; pp[1] = z; pp[2] = z;
    } else {
        for (long long c = c0; c < cells; ++c) {
            winner[c] = 0x7fffffff;
            mask[c] = 0;
            idx[3 * c] = -1; idx[3 * c + 1] = -1; idx[3 * c + 2] = -1;
            pts[3 * c] = 0.f; pts[3 * c + 1] = 0.f; pts[3 * c + 2] = 0.f;
        }
    }
}

// intersection.cu:8-27 -- note that corner 0 is never NaN-tested by the reference; a NaN there leaves
// min = max = NaN, after which neither `iso < min` nor `iso > max` is ever true: every isovalue hits.
__device__ __forceinline__ bool min_max8(const float4 lo, const float4 hi, float& mn, float& mx) {
    const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    mn = v[0]; mx = v[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        if (v[i] != v[i]) return false;
        if (v[i] < mn) mn = v[i];
        if (v[i] > mx) mx = v[i];
    }
    return true;
}

// window of sorted positions with !(iso < mn) && !(iso > mx)   (intersection.cu:98,107,115)
__device__ __forceinline__ void iso_window(const float* __restrict__ sorted, int n, float mn, float mx, int& lo, int& hi) {
    if (mn != mn) { lo = 0; hi = n; return; }
    int a = 0, b = n;
    while (a < b) { const int m = (a + b) >> 1; if (sorted[m] < mn) a = m + 1; else b = m; }
    lo = a;
    b = n;
    while (a < b) { const int m = (a + b) >> 1; if (!(sorted[m] > mx)) a = m + 1; else b = m; }
    hi = a;
}

struct VoxelFields {
    float4 s[2], a[2], b[2];
};

__device__ __forceinline__ void load_voxel(const float* __restrict__ sf, const float* __restrict__ af,
                                           const float* __restrict__ bf, long long v, VoxelFields& f) {
    const float4* sp = reinterpret_cast<const float4*>(sf) + 2 * v;
    const float4* ap = reinterpret_cast<const float4*>(af) + 2 * v;
    const float4* bp = reinterpret_cast<const float4*>(bf) + 2 * v;
    f.s[0] = __ldcs(sp); f.s[1] = __ldcs(sp + 1);
    f.a[0] = __ldcs(ap); f.a[1] = __ldcs(ap + 1);
    f.b[0] = __ldcs(bp); f.b[1] = __ldcs(bp + 1);
}

struct Windows {
    int s0, s1, a0, a1, b0, b1;
    bool any;
};

__device__ __forceinline__ Windows voxel_windows(const VoxelFields& f, const IsectWs& w, int, int, int) {
    const int Ns = w.hdr[4], N1 = w.hdr[5], N2 = w.hdr[6];  // finite isovalues only
    Windows r{};
    r.any = false;
    float amn, amx, bmn, bmx, smn, smx;
    if (!min_max8(f.a[0], f.a[1], amn, amx)) return r;  // order of intersection.cu:79-91
    if (!min_max8(f.b[0], f.b[1], bmn, bmx)) return r;
    if (!min_max8(f.s[0], f.s[1], smn, smx)) return r;
    iso_window(w.iso_sorted[1], N1, amn, amx, r.a0, r.a1);
    if (r.a0 >= r.a1) return r;
    iso_window(w.iso_sorted[2], N2, bmn, bmx, r.b0, r.b1);
    if (r.b0 >= r.b1) return r;
    iso_window(w.iso_sorted[0], Ns, smn, smx, r.s0, r.s1);
    if (r.s0 >= r.s1) return r;
    r.any = true;
    return r;
}

// pass 1: elect the winner voxel of every cell, compact the voxels that hit anything.
__global__ void __launch_bounds__(256)
isect_elect_kernel(const float* __restrict__ sf, const float* __restrict__ af, const float* __restrict__ bf,
                   long long voxels, int Ns, int N1, int N2, IsectWs w) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (v < voxels) {
        VoxelFields f;
        load_voxel(sf, af, bf, v, f);
        const Windows r = voxel_windows(f, w, Ns, N1, N2);
        if (r.any) {
            hit = true;
            for (int ja = r.a0; ja < r.a1; ++ja) {
                const int i1 = w.iso_perm[1][ja];
                for (int jb = r.b0; jb < r.b1; ++jb) {
                    const int i2 = w.iso_perm[2][jb];
                    for (int js = r.s0; js < r.s1; ++js) {
                        const int is = w.iso_perm[0][js];
                        const long long c = ((long long)is * N1 + i1) * N2 + i2;
                        atomicMin(&w.winner[c], (int)v);
                    }
                }
            }
        }
    }
    // warp-aggregated append to the work list
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(&w.hdr[0], __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (hit) w.worklist[base + __popc(m & ((1u << lane) - 1u))] = (int)v;
    }
}

__constant__ float c_off[8][3] = {  // intersection.cu:55-64
    {-0.5f, -0.5f, -0.5f}, {0.5f, -0.5f, -0.5f}, {-0.5f, 0.5f, -0.5f}, {0.5f, 0.5f, -0.5f},
    {-0.5f, -0.5f, 0.5f},  {0.5f, -0.5f, 0.5f},  {-0.5f, 0.5f, 0.5f},  {0.5f, 0.5f, 0.5f}};
__constant__ int c_edge[12][2] = {  // intersection.cu:67-71
    {0, 1}, {0, 2}, {1, 3}, {2, 3}, {4, 5}, {4, 6}, {5, 7}, {6, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

// intersection.cu:135-160: sum of the edge-crossing points of one field, inclusive edge test, unguarded division.
__device__ __forceinline__ int edge_sum(const float4 lo, const float4 hi, float iso, const float* ctr, float vs,
                                        float* acc) {
    const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    int cnt = 0;
    acc[0] = acc[1] = acc[2] = 0.f;
#pragma unroll
    for (int e = 0; e < 12; ++e) {
        const int c1 = c_edge[e][0], c2 = c_edge[e][1];
        const float a = v[c1], b = v[c2];
        if (iso >= fminf(a, b) && iso <= fmaxf(a, b)) {
            const float t = __fdiv_rn(iso - a, b - a);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float x1 = fmaf(vs, c_off[c1][d], ctr[d]);
                const float x2 = fmaf(vs, c_off[c2][d], ctr[d]);
                acc[d] += fmaf(t, x2 - x1, x1);
            }
            ++cnt;
        }
    }
    return cnt;
}

// pass 2: winners write their cells.  One thread per work-list voxel.
__global__ void __launch_bounds__(256)
isect_write_kernel(const float* __restrict__ sf, const float* __restrict__ af, const float* __restrict__ bf,
                   const float* __restrict__ centers, int Dx, int Ns, int N1, int N2, IsectWs w,
                   uint8_t* __restrict__ mask, int16_t* __restrict__ idx, float* __restrict__ pts) {
    const int nwork = w.hdr[0];
    const float vs = (float)(2.0 / (double)(float)Dx);  // intersection.cu:54
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nwork; i += gridDim.x * blockDim.x) {
        const long long v = w.worklist[i];
        VoxelFields f;
        load_voxel(sf, af, bf, v, f);
        const Windows r = voxel_windows(f, w, Ns, N1, N2);
        if (!r.any) continue;
        const float ctr[3] = {centers[3 * v], centers[3 * v + 1], centers[3 * v + 2]};
        for (int ja = r.a0; ja < r.a1; ++ja) {
            const int i1 = w.iso_perm[1][ja];
            const float aiso = w.iso_sorted[1][ja];
            float pa[3];
            const int ca = edge_sum(f.a[0], f.a[1], aiso, ctr, vs, pa);
            for (int jb = r.b0; jb < r.b1; ++jb) {
                const int i2 = w.iso_perm[2][jb];
                const float biso = w.iso_sorted[2][jb];
                float pb[3];
                const int cb = edge_sum(f.b[0], f.b[1], biso, ctr, vs, pb);
                for (int js = r.s0; js < r.s1; ++js) {
                    const int is = w.iso_perm[0][js];
                    const long long c = ((long long)is * N1 + i1) * N2 + i2;
                    if (w.winner[c] != (int)v) continue;
                    const float siso = w.iso_sorted[0][js];
                    float ps[3];
                    const int cs = edge_sum(f.s[0], f.s[1], siso, ctr, vs, ps);
                    idx[3 * c] = (int16_t)is; idx[3 * c + 1] = (int16_t)i1; idx[3 * c + 2] = (int16_t)i2;
                    mask[c] = 1;
#pragma unroll
                    for (int d = 0; d < 3; ++d) {  // intersection.cu:203-213
                        const float xs = cs > 0 ? __fdiv_rn(ps[d], (float)cs) : ps[d];
                        const float xa = ca > 0 ? __fdiv_rn(pa[d], (float)ca) : pa[d];
                        const float xb = cb > 0 ? __fdiv_rn(pb[d], (float)cb) : pb[d];
                        pts[3 * c + d] = (float)((double)(xs + xa + xb) / 3.0);
                    }
                }
            }
        }
    }
}

}  // namespace
}  // namespace inf3dp

using namespace inf3dp;

extern "C" {

size_t inf3dp_fields_intersection_workspace_bytes(int Dx, int Dy, int Dz, int Ns, int N1, int N2) {
    if (Dx < 0 || Dy < 0 || Dz < 0 || Ns < 0 || N1 < 0 || N2 < 0) return 0;
    const long long voxels = (long long)Dx * Dy * Dz, cells = (long long)Ns * N1 * N2;
    return carve_isect(nullptr, voxels > 0 ? voxels : 1, cells > 0 ? cells : 1, Ns, N1, N2).total_bytes;
}

int inf3dp_fields_intersection(const float* sf, const float* af, const float* bf, const float* s_iso,
                               const float* a_iso, const float* b_iso, const float* centers, int Dx, int Dy, int Dz,
                               int Ns, int N1, int N2, uint8_t* mask, int16_t* idx, float* pts, void* ws,
                               size_t ws_bytes, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(Dx >= 0 && Dy >= 0 && Dz >= 0 && Ns >= 0 && N1 >= 0 && N2 >= 0, "fields_intersection: negative size");
    const long long voxels = (long long)Dx * Dy * Dz, cells = (long long)Ns * N1 * N2;
    if (cells == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(voxels < 0x7fffffffll, "fields_intersection: %lld voxels exceed the int32 voxel index", voxels);
    INF3DP_CHECK_ARG(Ns <= 32767 && N1 <= 32767 && N2 <= 32767,
                     "fields_intersection: isovalue counts must fit the int16 index output (intersection.cu:247)");
    INF3DP_CHECK_ARG(mask && idx && pts && ws && s_iso && a_iso && b_iso, "fields_intersection: null pointer");
    INF3DP_CHECK_ARG(voxels == 0 || (sf && af && bf && centers), "fields_intersection: null pointer");
    INF3DP_CHECK_ARG(((reinterpret_cast<uintptr_t>(sf) | reinterpret_cast<uintptr_t>(af) | reinterpret_cast<uintptr_t>(bf) |
                       reinterpret_cast<uintptr_t>(pts) | reinterpret_cast<uintptr_t>(idx) |
                       reinterpret_cast<uintptr_t>(mask)) & 15) == 0,
                     "fields_intersection: field and output pointers must be 16-byte aligned");
    IsectWs w = carve_isect(ws, voxels > 0 ? voxels : 1, cells, Ns, N1, N2);
    if (ws_bytes < w.total_bytes) {
        set_error("fields_intersection: workspace has %zu bytes, need %zu", ws_bytes, w.total_bytes);
        return INF3DP_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    isect_prep_kernel<<<1, 1024, 0, st>>>(s_iso, a_iso, b_iso, Ns, N1, N2, w);
    INF3DP_LAUNCH_CHECK("isect_prep_kernel");
    const long long quads = (cells + 3) / 4;
    isect_init_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, st>>>(cells, w.winner, mask, idx, pts);
    INF3DP_LAUNCH_CHECK("isect_init_kernel");
    if (voxels > 0) {
        isect_elect_kernel<<<(unsigned)((voxels + 255) / 256), 256, 0, st>>>(sf, af, bf, voxels, Ns, N1, N2, w);
        INF3DP_LAUNCH_CHECK("isect_elect_kernel");
        isect_write_kernel<<<num_sms() * 8, 256, 0, st>>>(sf, af, bf, centers, Dx, Ns, N1, N2, w, mask, idx, pts);
        INF3DP_LAUNCH_CHECK("isect_write_kernel");
    }
    return INF3DP_OK;
}

}  // extern "C"
