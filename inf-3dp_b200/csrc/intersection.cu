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
//     elects the LOWEST linear voxel index per cell with atomicMin and pass 2 -- one thread per CELL -- evaluates
//     the winner voxel for that cell (or writes the defaults).  Deterministic, one full read of the fields.
//   * inf3dp_fields_intersection_grid reads the three [N,N,N] sample grids directly: the reference's host-side
//     8x corner expansion (toolpath_utils.py:986-1019) and voxel-centre tensor (:1021-1033) are never built
#include <algorithm>

#include "common.cuh"

namespace inf3dp {
namespace {

struct IsectWs {
    int* hdr;         // [4..6] finite isovalue counts (slice, l1, l2)
    float* iso_sorted[3];
    int* iso_perm[3];
    int* winner;      // [cells]
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
    (void)voxels;
    w.total_bytes = o;
    return w;
}

// Rank sort (ties keep input order), one WARP per element: the lanes split the comparison loop and reduce with
// shuffles (a thread per element made the prep kernel 20 us long at Ns = 500).  NaN isovalues are parked behind the
// finite ones and never searched: a NaN isovalue is ignored here (the reference would let it hit every voxel,
// intersection.cu:98 -- a documented divergence on meaningless input).
__device__ __forceinline__ void rank_sort(const float* __restrict__ iso, int n, float* sorted, int* perm, int* n_finite) {
    const int lane = threadIdx.x & 31, nwarps = (blockDim.x >> 5) * gridDim.x;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (int i = warp; i < n; i += nwarps) {
        const float vi = iso[i];
        int less = 0, finite = 0, nan_before = 0;
        for (int j = lane; j < n; j += 32) {
            const float vj = iso[j];
            less += (vj < vi) || (vj == vi && j < i);
            finite += (vj == vj);
            nan_before += (vj != vj) && j < i;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            less += __shfl_xor_sync(0xffffffffu, less, o);
            finite += __shfl_xor_sync(0xffffffffu, finite, o);
            nan_before += __shfl_xor_sync(0xffffffffu, nan_before, o);
        }
        if (lane == 0) {
            const int rank = (vi == vi) ? less : finite + nan_before;
            sorted[rank] = vi;
            perm[rank] = i;
            if (i == 0) *n_finite = finite;
        }
    }
    if (n == 0 && threadIdx.x == 0 && blockIdx.x == 0) *n_finite = 0;
}

// grid (8, 3): blockIdx.y = isovalue array, its 8 blocks split the elements (each stages the array in shared memory).
__global__ void isect_prep_kernel(const float* s_iso, const float* a_iso, const float* b_iso, int Ns, int N1, int N2,
                                  IsectWs w) {
    extern __shared__ float s_prep[];
    const int which = blockIdx.y;
    const float* iso = which == 0 ? s_iso : (which == 1 ? a_iso : b_iso);
    const int n = which == 0 ? Ns : (which == 1 ? N1 : N2);
    if (which == 0 && blockIdx.x == 0 && threadIdx.x < 4) w.hdr[threadIdx.x] = 0;
    const bool staged = n <= 8192;
    if (staged) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) s_prep[i] = iso[i];
        __syncthreads();
    }
    rank_sort(staged ? s_prep : iso, n, w.iso_sorted[which], w.iso_perm[which], &w.hdr[4 + which]);   // blocks split the elements
}

// Winner initialisation (the outputs are initialised by the cell kernel).
__global__ void isect_init_kernel(long long cells, int* __restrict__ winner) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // quad of cells
    const long long c0 = q * 4;
    if (c0 >= cells) return;
    if (c0 + 4 <= cells) *reinterpret_cast<int4*>(winner + c0) = make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff);
    else for (long long c = c0; c < cells; ++c) winner[c] = 0x7fffffff;
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

// First sorted position p with !PRED(sorted[p]) where PRED is monotone (true ... true false ... false): a lower / upper
// bound.  Isovalue sets are almost always evenly spaced (np.linspace, toolpath_utils.py:223), so the search starts at the
// position linear interpolation predicts and gallops from there: 2-3 probes instead of log2(n) for such sets, never more
// than ~2 log2(n) for arbitrary ones; the result is exactly the binary search's.
template <typename Pred>
__device__ __forceinline__ int bound_from_guess(const float* __restrict__ sorted, int n, int guess, Pred pred) {
    if (n <= 0) return 0;
    int g = min(max(guess, 0), n - 1);
    int lo, hi;                                   // invariant: pred holds on [.., lo), fails on [hi, ..)
    if (pred(sorted[g])) {
        lo = g + 1;
        int step = 1;
        hi = n;
        while (lo < n) {
            const int probe = min(lo + step - 1, n - 1);
            if (!pred(sorted[probe])) { hi = probe; break; }
            lo = probe + 1;
            step <<= 1;
        }
        if (lo >= n) return n;
    } else {
        hi = g;
        int step = 1;
        lo = 0;
        while (hi > 0) {
            const int probe = max(hi - step, 0);
            if (pred(sorted[probe])) { lo = probe + 1; break; }
            hi = probe;
            step <<= 1;
        }
        if (hi <= 0) return 0;
    }
    while (lo < hi) { const int m = (lo + hi) >> 1; if (pred(sorted[m])) lo = m + 1; else hi = m; }
    return lo;
}

// window of sorted positions with !(iso < mn) && !(iso > mx)   (intersection.cu:98,107,115)
// first / scale: sorted[0] and (n - 1) / (sorted[n-1] - sorted[0]) (0 when degenerate) for the interpolation guess.
//
// `uniform` (established per block by load_tables): the map f(v) = fl(fl(v - first) * scale) sends every table entry j to
// within kUniTol of j.  f is monotone (each float operation is), so for any bound value m with p = f(m): every entry
// j < p - kUniTol is < m and every entry j > p + kUniTol is > m.  Unless an integer lies within kUniTol of p, the lower bound
// (first entry >= m) and the upper bound (first entry > m) are both ceil(p) -- EXACTLY, whatever the rounding of f -- and no
// table probe is needed; the ~0.2 % of bounds that sit next to an integer take the probing search.  np.linspace isovalue
// sets (toolpath_utils.py:223) are uniform to ~1e-4; any other set fails the check and always takes the search.
constexpr float kUniTol = 1e-3f;

__device__ __forceinline__ void iso_window(const float* __restrict__ sorted, int n, float first, float scale, bool uniform,
                                           float mn, float mx, int& lo, int& hi) {
    if (mn != mn) { lo = 0; hi = n; return; }
    const float gl = (mn - first) * scale, gh = (mx - first) * scale;
    const bool fast_lo = uniform && fabsf(gl - rintf(gl)) > kUniTol, fast_hi = uniform && fabsf(gh - rintf(gh)) > kUniTol;
    // float -> int conversion saturates and maps NaN to 0: any value is a valid starting point
    if (fast_lo) lo = min(max(__float2int_ru(gl), 0), n);
    else lo = bound_from_guess(sorted, n, __float2int_ru(gl), [mn](float v) { return v < mn; });
    if (fast_hi) hi = min(max(__float2int_ru(gh), 0), n);
    else hi = bound_from_guess(sorted, n, __float2int_rd(gh) + 1, [mx](float v) { return !(v > mx); });
    if (hi < lo) hi = lo;
}

// The three sorted isovalue arrays and their permutations, staged in shared memory per block (a voxel makes ~25
// dependent probes into them; from global memory those probes were the elect kernel's critical path).
struct IsoTables {
    const float* sorted[3];
    const int* perm[3];
    int n[3];   // finite isovalues only
    float first[3], scale[3];   // interpolation guess: position ~ (value - first) * scale
    bool uniform[3];            // every entry j maps to within kUniTol of j (see iso_window)
};
constexpr int kMaxTableEntries = 6000;   // 48 KB of shared memory; larger isovalue sets are probed in global memory

__device__ __forceinline__ IsoTables load_tables(const IsectWs& w, unsigned char* smem, int Ns, int N1, int N2) {
    IsoTables t;
    const int full[3] = {Ns, N1, N2};
    t.n[0] = w.hdr[4]; t.n[1] = w.hdr[5]; t.n[2] = w.hdr[6];
    for (int i = 0; i < 3; ++i) {
        const float a = t.n[i] > 0 ? w.iso_sorted[i][0] : 0.f, b = t.n[i] > 0 ? w.iso_sorted[i][t.n[i] - 1] : 0.f;
        t.first[i] = a;
        t.scale[i] = (t.n[i] > 1 && b > a) ? __fdividef((float)(t.n[i] - 1), b - a) : 0.f;
        if (!(fabsf(t.scale[i]) <= 3.402823466e38f)) t.scale[i] = 0.f;
    }
    if (Ns + N1 + N2 > kMaxTableEntries) {
        for (int i = 0; i < 3; ++i) { t.sorted[i] = w.iso_sorted[i]; t.perm[i] = w.iso_perm[i]; }
    } else {
        float* fs = reinterpret_cast<float*>(smem);
        int* ps = reinterpret_cast<int*>(fs + (Ns + N1 + N2));
        int o = 0;
        for (int i = 0; i < 3; ++i) {
            for (int j = threadIdx.x; j < full[i]; j += blockDim.x) { fs[o + j] = w.iso_sorted[i][j]; ps[o + j] = w.iso_perm[i][j]; }
            t.sorted[i] = fs + o;
            t.perm[i] = ps + o;
            o += full[i];
        }
        __syncthreads();
    }
    for (int i = 0; i < 3; ++i) {   // uniformity of the finite entries under the SAME map the window search applies
        bool ok = t.n[i] > 1 && t.scale[i] > 0.f;
        for (int j = threadIdx.x; j < t.n[i] && ok; j += blockDim.x)
            ok = fabsf((t.sorted[i][j] - t.first[i]) * t.scale[i] - (float)j) <= kUniTol;
        t.uniform[i] = __syncthreads_and(ok) != 0;
    }
    return t;
}

struct VoxelFields {
    float4 s[2], a[2], b[2];
};

// Where the corner values of a voxel come from.  GRID = false: the reference's corner-expanded [Dx,Dy,Dz,8] tensors
// (32 B per voxel and field) and its [Dx,Dy,Dz,3] voxel centres.  GRID = true: the [Dx+1,Dy+1,Dz+1] sample grids
// themselves (4 B per voxel and field from HBM, neighbours through L1/L2) -- the expansion of
// toolpath_utils.py:986-1019 (corner c = offsets (c&1, (c>>1)&1, (c>>2)&1) on the three grid axes) and the centres of
// toolpath_utils.py:1021-1033 are evaluated on the fly, bit-identically.
struct FieldSrc {
    const float *sf, *af, *bf;
    const float* centers;   // expanded mode only
    int Dy, Dz;             // voxel counts on axes 1, 2
    long long sx, sy;       // grid strides of axes 0, 1 (axis 2 is contiguous)
    float center_vs;        // (float)(2.0 / (Gx - 1)): get_voxel_centers' voxel size as torch applies it to a float32 tensor
};

__device__ __forceinline__ void voxel_xyz(const FieldSrc& src, unsigned v, int& x, int& y, int& z) {
    const unsigned u = v / (unsigned)src.Dz;    // voxels < 2^31: 32-bit index arithmetic
    z = (int)(v - u * (unsigned)src.Dz);
    x = (int)(u / (unsigned)src.Dy);
    y = (int)(u - (unsigned)x * (unsigned)src.Dy);
}

__device__ __forceinline__ void load_plane(const FieldSrc& src, long long b, float4& s, float4& a, float4& bb) {
    const long long o1 = b + src.sx, o2 = b + src.sy, o3 = o1 + src.sy;
    s = make_float4(__ldg(src.sf + b), __ldg(src.sf + o1), __ldg(src.sf + o2), __ldg(src.sf + o3));
    a = make_float4(__ldg(src.af + b), __ldg(src.af + o1), __ldg(src.af + o2), __ldg(src.af + o3));
    bb = make_float4(__ldg(src.bf + b), __ldg(src.bf + o1), __ldg(src.bf + o2), __ldg(src.bf + o3));
}

// one voxel by index (gather): expanded tensors or grids
template <bool GRID, bool STREAM>
__device__ __forceinline__ void load_voxel(const FieldSrc& src, long long v, VoxelFields& f) {
    if (GRID) {
        int x, y, z;
        voxel_xyz(src, (unsigned)v, x, y, z);
        const long long b = x * src.sx + y * src.sy + z;
        load_plane(src, b, f.s[0], f.a[0], f.b[0]);
        load_plane(src, b + 1, f.s[1], f.a[1], f.b[1]);
    } else {
        const float4* sp = reinterpret_cast<const float4*>(src.sf) + 2 * v;
        const float4* ap = reinterpret_cast<const float4*>(src.af) + 2 * v;
        const float4* bp = reinterpret_cast<const float4*>(src.bf) + 2 * v;
        if (STREAM) {
            f.s[0] = __ldcs(sp); f.s[1] = __ldcs(sp + 1);
            f.a[0] = __ldcs(ap); f.a[1] = __ldcs(ap + 1);
            f.b[0] = __ldcs(bp); f.b[1] = __ldcs(bp + 1);
        } else {
            f.s[0] = __ldg(sp); f.s[1] = __ldg(sp + 1);
            f.a[0] = __ldg(ap); f.a[1] = __ldg(ap + 1);
            f.b[0] = __ldg(bp); f.b[1] = __ldg(bp + 1);
        }
    }
}

__device__ __forceinline__ float4 shfl_down4(const float4 v) {
    return make_float4(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1),
                       __shfl_down_sync(0xffffffffu, v.z, 1), __shfl_down_sync(0xffffffffu, v.w, 1));
}

// grids, consecutive lanes = consecutive voxels (called by all 32 lanes): a lane loads the four corners of its lower z
// plane per field; its upper plane is the next lane's lower plane (shuffle) unless the row or the warp ends there
__device__ __forceinline__ void load_voxel_grid_warp(const FieldSrc& src, long long v, bool active, VoxelFields& f) {
    int x = 0, y = 0, z = 0;
    long long b = 0;
    f.s[0] = f.a[0] = f.b[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
        voxel_xyz(src, (unsigned)v, x, y, z);
        b = x * src.sx + y * src.sy + z;
        load_plane(src, b, f.s[0], f.a[0], f.b[0]);
    }
    f.s[1] = shfl_down4(f.s[0]);
    f.a[1] = shfl_down4(f.a[0]);
    f.b[1] = shfl_down4(f.b[0]);
    if (active && ((threadIdx.x & 31) == 31 || z + 1 == src.Dz)) load_plane(src, b + 1, f.s[1], f.a[1], f.b[1]);
}

template <bool GRID>
__device__ __forceinline__ void load_center(const FieldSrc& src, long long v, float* ctr) {
    if (GRID) {   // (index + 0.5) * voxel_size + (-1), each step rounded to float32 as torch does
        int x, y, z;
        voxel_xyz(src, (unsigned)v, x, y, z);
        ctr[0] = __fadd_rn(__fmul_rn((float)x + 0.5f, src.center_vs), -1.0f);
        ctr[1] = __fadd_rn(__fmul_rn((float)y + 0.5f, src.center_vs), -1.0f);
        ctr[2] = __fadd_rn(__fmul_rn((float)z + 0.5f, src.center_vs), -1.0f);
    } else {
        ctr[0] = __ldg(src.centers + 3ll * v); ctr[1] = __ldg(src.centers + 3ll * v + 1); ctr[2] = __ldg(src.centers + 3ll * v + 2);
    }
}

struct Windows {
    int s0, s1, a0, a1, b0, b1;
    bool any;
};

__device__ __forceinline__ Windows voxel_windows(const VoxelFields& f, const IsoTables& w) {
    const int Ns = w.n[0], N1 = w.n[1], N2 = w.n[2];  // finite isovalues only
    Windows r{};
    r.any = false;
    float amn, amx, bmn, bmx, smn, smx;
    if (!min_max8(f.a[0], f.a[1], amn, amx)) return r;  // order of intersection.cu:79-91
    if (!min_max8(f.b[0], f.b[1], bmn, bmx)) return r;
    if (!min_max8(f.s[0], f.s[1], smn, smx)) return r;
    iso_window(w.sorted[1], N1, w.first[1], w.scale[1], w.uniform[1], amn, amx, r.a0, r.a1);
    if (r.a0 >= r.a1) return r;
    iso_window(w.sorted[2], N2, w.first[2], w.scale[2], w.uniform[2], bmn, bmx, r.b0, r.b1);
    if (r.b0 >= r.b1) return r;
    iso_window(w.sorted[0], Ns, w.first[0], w.scale[0], w.uniform[0], smn, smx, r.s0, r.s1);
    if (r.s0 >= r.s1) return r;
    r.any = true;
    return r;
}

// pass 1: elect the winner voxel of every cell, compact the voxels that hit anything.
template <bool GRID>
__global__ void __launch_bounds__(256)
isect_elect_kernel(const FieldSrc src, long long voxels, int Ns, int N1, int N2, IsectWs w) {
    extern __shared__ unsigned char s_tab[];
    const IsoTables tb = load_tables(w, s_tab, Ns, N1, N2);
    // persistent blocks: the tables are staged once per block, voxels are visited in block-sized strides
    for (long long base = (long long)blockIdx.x * blockDim.x; base < voxels; base += (long long)gridDim.x * blockDim.x) {
    const long long v = base + threadIdx.x;
    bool hit = false;
    VoxelFields f;
    if (GRID) load_voxel_grid_warp(src, v, v < voxels, f);
    if (v < voxels) {
        if (!GRID) load_voxel<false, false>(src, v, f);   // cached loads: the two 16-byte halves of a 32-byte row share their sector in L1
        const Windows r = voxel_windows(f, tb);
        if (r.any) {
            hit = true;
            for (int ja = r.a0; ja < r.a1; ++ja) {
                const int i1 = tb.perm[1][ja];
                for (int jb = r.b0; jb < r.b1; ++jb) {
                    const int i2 = tb.perm[2][jb];
                    for (int js = r.s0; js < r.s1; ++js) {
                        const int is = tb.perm[0][js];
                        const long long c = ((long long)is * N1 + i1) * N2 + i2;
                        // voxels are visited in roughly ascending order and ~3 of them hit every cell: most of the time
                        // the cell already holds a smaller index, and a plain load is much cheaper than an L2 atomic
                        if (__ldcg(&w.winner[c]) > (int)v) atomicMin(&w.winner[c], (int)v);
                    }
                }
            }
        }
    }
    (void)hit;
    }
}

// intersection.cu:55-71: corner c sits at centre + voxel_size * (+-0.5 per axis, bit d of c = axis d), the 12 edges are
// (0,1) (0,2) (1,3) (2,3) (4,5) (4,6) (5,7) (6,7) (0,4) (1,5) (2,6) (3,7) -- packed one nibble per edge.
constexpr unsigned long long kEdgeC1 = 0x321065442100ull;   // first corner of edge e in bits [4e, 4e+4)
constexpr unsigned long long kEdgeC2 = 0x765477653321ull;   // second corner

// Mean edge-crossing point of one field at one isovalue (intersection.cu:135-160 + the division of :203-209): the 12
// edges in the reference's order with compile-time corner ids (the corner values stay in registers: no staging, no dynamic
// indexing), inclusive test, unguarded IEEE division, the same FMA contraction nvcc applies to
// `coord1 + t * (coord2 - coord1)`.  On an axis the edge does not run along, coord2 - coord1 is exactly 0, so the term is
// coord1 -- or NaN when t is not finite (flat edge exactly at the isovalue), which fmaf(t, 0, x) preserves.
__device__ __forceinline__ void edge_mean(const float4 q0, const float4 q1, float iso, const float* lo, const float* hi,
                                          float* out) {
    const float v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    int cnt = 0;
    const float d0 = hi[0] - lo[0], d1 = hi[1] - lo[1], d2 = hi[2] - lo[2];
#pragma unroll
    for (int e = 0; e < 12; ++e) {
        const int c1 = (int)((kEdgeC1 >> (4 * e)) & 15), c2 = (int)((kEdgeC2 >> (4 * e)) & 15);
        const float a = v[c1], b = v[c2];
        if (iso >= fminf(a, b) && iso <= fmaxf(a, b)) {
            const float t = __fdiv_rn(iso - a, b - a);
            const int ax = (c1 ^ c2) == 1 ? 0 : ((c1 ^ c2) == 2 ? 1 : 2);     // compile time: the axis this edge runs along
            const float x1 = (c1 & 1) ? hi[0] : lo[0], y1 = (c1 & 2) ? hi[1] : lo[1], z1 = (c1 & 4) ? hi[2] : lo[2];
            acc0 += fmaf(t, ax == 0 ? d0 : 0.f, x1);
            acc1 += fmaf(t, ax == 1 ? d1 : 0.f, y1);
            acc2 += fmaf(t, ax == 2 ? d2 : 0.f, z1);
            ++cnt;
        }
    }
    if (cnt > 0) {
        const float fc = (float)cnt;
        out[0] = __fdiv_rn(acc0, fc); out[1] = __fdiv_rn(acc1, fc); out[2] = __fdiv_rn(acc2, fc);
    } else {
        out[0] = acc0; out[1] = acc1; out[2] = acc2;
    }
}

// pass 2, cell centric: one thread per cell.  A cell without a winner gets the "nothing here" outputs (mask 0,
// indices -1, point 0: no separate initialisation pass); otherwise the cell's winner voxel is re-read (96 B of corner
// fields + its centre) and the three edge sums are evaluated for exactly this cell's isovalues.  Uniform work per
// thread, no index windows.  A block covers 8 consecutive slice isovalues x 32 consecutive lattice-2 isovalues of one
// lattice-1 isovalue: every warp writes 32 consecutive cells (coalesced rows), and the 8 warps mostly re-read the SAME
// winner voxels (a voxel typically spans 2-3 neighbouring slice isovalues), which the L1 serves.
constexpr int kCellTileS = 8, kCellTile2 = 32;

template <bool GRID>
__global__ void __launch_bounds__(256)
isect_cell_kernel(const FieldSrc src, const float* __restrict__ s_iso, const float* __restrict__ a_iso,
                  const float* __restrict__ b_iso, int Dx, int Ns, int N1, int N2, long long tiles, IsectWs w,
                  uint8_t* __restrict__ mask, int16_t* __restrict__ idx, float* __restrict__ pts) {
    const float vs = (float)(2.0 / (double)(float)Dx);  // intersection.cu:54
    const int tiles2 = (N2 + kCellTile2 - 1) / kCellTile2;
    const int r = threadIdx.x >> 5, q = threadIdx.x & 31;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int t2 = (int)(t % tiles2);
        const long long u = t / tiles2;
        const int i1 = (int)(u % N1);
        const int ts = (int)(u / N1);
        const int is = ts * kCellTileS + r, i2 = t2 * kCellTile2 + q;
        if (is >= Ns || i2 >= N2) continue;
        const long long c = ((long long)is * N1 + i1) * N2 + i2;
        const int v = __ldcs(&w.winner[c]);
        if (v == 0x7fffffff) {
            mask[c] = 0;
            idx[3 * c] = -1; idx[3 * c + 1] = -1; idx[3 * c + 2] = -1;
            pts[3 * c] = 0.f; pts[3 * c + 1] = 0.f; pts[3 * c + 2] = 0.f;
            continue;
        }
        VoxelFields f;
        load_voxel<GRID, false>(src, v, f);
        float ctr[3];
        load_center<GRID>(src, v, ctr);
        // corner coordinates: centre + voxel_size * (-0.5 | +0.5), the two values per axis shared by all three fields
        float lo[3], hi[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) { lo[d] = fmaf(vs, -0.5f, ctr[d]); hi[d] = fmaf(vs, 0.5f, ctr[d]); }
        float pa[3], pb[3], ps[3];
        edge_mean(f.a[0], f.a[1], a_iso[i1], lo, hi, pa);
        edge_mean(f.b[0], f.b[1], b_iso[i2], lo, hi, pb);
        edge_mean(f.s[0], f.s[1], s_iso[is], lo, hi, ps);
        idx[3 * c] = (int16_t)is; idx[3 * c + 1] = (int16_t)i1; idx[3 * c + 2] = (int16_t)i2;
        mask[c] = 1;
#pragma unroll
        // intersection.cu:203-213 divides by the double literal 3.0: float(double(v) / 3.0).  That equals the correctly rounded
        // single-precision quotient: v / 3 has a periodic binary expansion (or is exact), so the double result never sits
        // within 2^-53 of a single-precision rounding boundary and the second rounding cannot change the outcome.
        for (int d = 0; d < 3; ++d) pts[3 * c + d] = __fdiv_rn(ps[d] + pa[d] + pb[d], 3.0f);
    }
}


template <bool GRID>
int isect_run(const FieldSrc& src, const float* s_iso, const float* a_iso, const float* b_iso, int Dx, int Dy, int Dz,
              int Ns, int N1, int N2, uint8_t* mask, int16_t* idx, float* pts, void* ws, size_t ws_bytes,
              inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(Dx >= 0 && Dy >= 0 && Dz >= 0 && Ns >= 0 && N1 >= 0 && N2 >= 0, "fields_intersection: negative size");
    const long long voxels = (long long)Dx * Dy * Dz, cells = (long long)Ns * N1 * N2;
    if (cells == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(voxels < 0x7fffffffll, "fields_intersection: %lld voxels exceed the int32 voxel index", voxels);
    INF3DP_CHECK_ARG(Ns <= 32767 && N1 <= 32767 && N2 <= 32767,
                     "fields_intersection: isovalue counts must fit the int16 index output (intersection.cu:247)");
    INF3DP_CHECK_ARG(mask && idx && pts && ws && s_iso && a_iso && b_iso, "fields_intersection: null pointer");
    INF3DP_CHECK_ARG(voxels == 0 || (src.sf && src.af && src.bf), "fields_intersection: null pointer");
    INF3DP_CHECK_ARG(((reinterpret_cast<uintptr_t>(pts) | reinterpret_cast<uintptr_t>(idx) |
                       reinterpret_cast<uintptr_t>(mask)) & 15) == 0,
                     "fields_intersection: field and output pointers must be 16-byte aligned");
    IsectWs w = carve_isect(ws, voxels > 0 ? voxels : 1, cells, Ns, N1, N2);
    if (ws_bytes < w.total_bytes) {
        set_error("fields_intersection: workspace has %zu bytes, need %zu", ws_bytes, w.total_bytes);
        return INF3DP_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    isect_prep_kernel<<<dim3(8, 3), 1024, 8192 * sizeof(float), st>>>(s_iso, a_iso, b_iso, Ns, N1, N2, w);
    INF3DP_LAUNCH_CHECK("isect_prep_kernel");
    const long long quads = (cells + 3) / 4;
    isect_init_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, st>>>(cells, w.winner);
    INF3DP_LAUNCH_CHECK("isect_init_kernel");
    if (voxels > 0) {
        const size_t tab_smem = (Ns + N1 + N2) <= kMaxTableEntries ? (size_t)(Ns + N1 + N2) * 8 : 0;
        if (tab_smem > 40 * 1024) {
            INF3DP_CUDA(cudaFuncSetAttribute(isect_elect_kernel<GRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_smem));
        }
        isect_elect_kernel<GRID><<<(unsigned)std::min<long long>((voxels + 255) / 256, (long long)num_sms() * 16), 256, tab_smem, st>>>(src, voxels, Ns, N1, N2, w);
        INF3DP_LAUNCH_CHECK("isect_elect_kernel");
    }
    {   // cells without a winner (all of them when there are no voxels) get the default outputs here
        const long long tiles = (long long)((Ns + kCellTileS - 1) / kCellTileS) * N1 * ((N2 + kCellTile2 - 1) / kCellTile2);
        const long long blocks = std::min<long long>(tiles, (long long)num_sms() * 32);
        isect_cell_kernel<GRID><<<(unsigned)blocks, 256, 0, st>>>(src, s_iso, a_iso, b_iso, Dx, Ns, N1, N2, tiles, w, mask, idx, pts);
        INF3DP_LAUNCH_CHECK("isect_cell_kernel");
    }
    return INF3DP_OK;
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
    const long long voxels = (long long)Dx * Dy * Dz;
    INF3DP_CHECK_ARG(Dx >= 0 && Dy >= 0 && Dz >= 0, "fields_intersection: negative size");
    INF3DP_CHECK_ARG(voxels == 0 || centers, "fields_intersection: null pointer");
    INF3DP_CHECK_ARG(((reinterpret_cast<uintptr_t>(sf) | reinterpret_cast<uintptr_t>(af) | reinterpret_cast<uintptr_t>(bf)) & 15) == 0,
                     "fields_intersection: field and output pointers must be 16-byte aligned");
    FieldSrc src{sf, af, bf, centers, Dy, Dz, 0, 0, 0.f};
    return isect_run<false>(src, s_iso, a_iso, b_iso, Dx, Dy, Dz, Ns, N1, N2, mask, idx, pts, ws, ws_bytes, stream);
}

int inf3dp_fields_intersection_grid(const float* sg, const float* ag, const float* bg, const float* s_iso,
                                    const float* a_iso, const float* b_iso, int Gx, int Gy, int Gz, int Ns, int N1,
                                    int N2, uint8_t* mask, int16_t* idx, float* pts, void* ws, size_t ws_bytes,
                                    inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(Gx >= 1 && Gy >= 1 && Gz >= 1, "fields_intersection_grid: grids need at least one sample per axis");
    const int Dx = Gx - 1, Dy = Gy - 1, Dz = Gz - 1;
    FieldSrc src{sg, ag, bg, nullptr, Dy, Dz, (long long)Gy * Gz, (long long)Gz,
                 Gx > 1 ? (float)(2.0 / (double)(Gx - 1)) : 0.f};
    return isect_run<true>(src, s_iso, a_iso, b_iso, Dx, Dy, Dz, Ns, N1, N2, mask, idx, pts, ws, ws_bytes, stream);
}

}  // extern "C"
