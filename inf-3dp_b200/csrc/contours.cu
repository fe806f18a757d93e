// Isocontour extraction (marching triangles + polyline chaining) for sm_100a.
//
// Replaces contour_cuda/ext.cpp:5-17 -> isocontours.cu:216-293 of the reference:
//   extract_contours_kernel (:16-93)  one thread per face looping over ALL K isovalues, 36*K*F bytes of scratch
//   merge_segments_kernel   (:99-213) ONE thread per isovalue, O(S^2) global-memory scans
// with an output-identical pipeline built for the B200 memory system:
//
//   prep        sort the K isovalues once (rank sort) so a face only visits isovalues inside (fmin, fmax)
//   count/emit  one thread per face, binary search of its isovalue window, the reference's exact strict-sign
//               edge test; segments are slotted per isovalue by warp-aggregated atomics (match_any + one
//               atomicAdd per distinct isovalue per warp) into a compact O(sum S_k) pool -- the dense
//               [K,F,6] scratch of the reference is never materialised
//   hash/link   every segment endpoint lies on a mesh edge; the two triangles sharing that edge meet through a
//               (isovalue, edge) keyed open-addressing table in L2, then keep the link only if the two
//               independently interpolated points agree within the reference's 1e-6 merge radius (:111,178,188)
//   walk        one CTA per isovalue: the CTA pulls the isovalue's links/points through L1 and one thread walks
//               the chain in O(S) dependent loads (the reference is O(S^2)); seeds are taken in ascending face
//               order, which is the reference's result when its atomicAdd slots faces in index order
//   fill        the dense zero-padded [K,F,3] contract is written exactly once: the bulk zero fill (rows >= 2 S_k)
//               runs on a side stream concurrently with hash/link/walk, the remainder after the walk
//
// Output contract (identical to the reference): contours_merged [K,F,3] zero padded, per isovalue the row stream
// [seed.p1, far endpoint of each chained segment ...] (-1e9 row) ... ; contour_nums [K].
#include <mutex>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace inf3dp {
namespace {

constexpr int kFaceThreads = 256;
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

// big walk tier (multi-million-face meshes): list capacities per isovalue, kept in the workspace
constexpr int kBigMaxLoops = 2048, kBigMaxPaths = 1024;
constexpr int kWalkListInts = 7 * kBigMaxLoops + 3 * kBigMaxPaths;
constexpr int kWalkBigMinFaces = 1500000;   // pools at least this large carry the lists (see launch_walk)

struct ContourWs {
    // header (device): [0] overflow flag, [1] total segments, [2] hash slots in use (pow2), [3] segment capacity
    int* hdr;
    float* iso_sorted;   // [K]
    int* iso_perm;       // [K] sorted position -> original index
    int* seg_count;      // [K]
    int* seg_offset;     // [K+1]
    int* cursor;         // [K]
    int* rows_used;      // [K]
    int* seg_face;       // [cap]
    float* seg_pts;      // [cap][6]
    int* seg_slot;       // [cap][2] hash slot of each endpoint
    int* seg_link;       // [cap][2] partner endpoint id (2*seg+side) or -1
    int* seg_order;      // [2*cap] row -> endpoint id (or -2 separator), per isovalue at 2*seg_offset
    unsigned char* used; // [cap]
    unsigned long long* hkeys;  // [hcap]
    int* hvals;          // [hcap][2]
    int* walk_lists;     // [K][kWalkListInts] piece / chain lists of the big walk tier (null for small pools)
    int cap;
    int hcap;
    int defer;           // 1: the walk records the row order (endpoint ids) in seg_order, rows are written by the scatter kernel
    int compact;         // 1: rows of isovalue k go to merged + 6 * seg_offset[k] (at most 2 S_k rows), no dense contract
    size_t total_bytes;
};

// Default segment pool: one segment per face on average (measured meshes need ~0.3-0.5), never more than the
// worst case K*F.  A caller that gets INF3DP_EOVERFLOW passes a larger workspace (see *_workspace_bytes_segments).
inline int seg_capacity_default(int nF, int K) {
    long long c = (long long)nF + 65536;
    long long worst = (long long)nF * (long long)K;
    if (c > worst) c = worst;
    if (c < 1024) c = 1024;
    if (c > 0x3fffffff) c = 0x3fffffff;
    return (int)c;
}
inline int pow2_at_least(long long v) {
    long long p = 1024;
    while (p < v) p <<= 1;
    return (int)p;
}

ContourWs carve(void* base, int K, int cap) {
    ContourWs w{};
    w.cap = cap;
    w.hcap = pow2_at_least(2ll * w.cap);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    char* b = (char*)base;
    w.hdr = (int*)(b + take(64));
    w.iso_sorted = (float*)(b + take((size_t)K * 4));
    w.iso_perm = (int*)(b + take((size_t)K * 4));
    w.seg_count = (int*)(b + take((size_t)K * 4));
    w.seg_offset = (int*)(b + take((size_t)(K + 1) * 4));
    w.cursor = (int*)(b + take((size_t)K * 4));
    w.rows_used = (int*)(b + take((size_t)K * 4));
    w.seg_face = (int*)(b + take((size_t)w.cap * 4));
    w.seg_pts = (float*)(b + take((size_t)w.cap * 24));
    w.seg_slot = (int*)(b + take((size_t)w.cap * 8));
    w.seg_link = (int*)(b + take((size_t)w.cap * 8));
    w.seg_order = (int*)(b + take((size_t)w.cap * 8));
    w.used = (unsigned char*)(b + take((size_t)w.cap));
    w.hkeys = (unsigned long long*)(b + take((size_t)w.hcap * 8));
    w.hvals = (int*)(b + take((size_t)w.hcap * 8));
    w.walk_lists = w.cap >= kWalkBigMinFaces ? (int*)(b + take((size_t)K * kWalkListInts * 4)) : nullptr;
    w.total_bytes = o;
    return w;
}

// ---------------------------------------------------------------------------------------------------------
// isocontours.cu:9-11.  The product feeds a division, so there is no mul+add pair for -fmad to contract:
// every step is a correctly rounded IEEE operation, identical to the reference binary and to the CPU oracle.
__device__ __forceinline__ float interp_iso(float v0, float v1, float f0, float f1, float iso) {
    return __fadd_rn(v0, __fdiv_rn(__fmul_rn(__fsub_rn(iso, f0), __fsub_rn(v1, v0)), __fsub_rn(f1, f0)));
}

struct FaceData {
    int v[3];
    float f[3];
};

// isocontours.cu:42-92 for one (face, isovalue).  Returns true when exactly two edges cross; p receives the
// two points in edge order (v0v1, v1v2, v2v0), e0/e1 the local edge index of each.
__device__ __forceinline__ bool face_segment(const FaceData& fd, float iso, const float* __restrict__ V, float* p,
                                             int& e0, int& e1) {
    const float f0 = fd.f[0], f1 = fd.f[1], f2 = fd.f[2];
    const bool below = (f0 <= iso) && (f1 <= iso) && (f2 <= iso);
    const bool above = (f0 >= iso) && (f1 >= iso) && (f2 >= iso);
    if (below || above) return false;
    const bool c0 = __fmul_rn(__fsub_rn(f0, iso), __fsub_rn(f1, iso)) < 0.f;
    const bool c1 = __fmul_rn(__fsub_rn(f1, iso), __fsub_rn(f2, iso)) < 0.f;
    const bool c2 = __fmul_rn(__fsub_rn(f2, iso), __fsub_rn(f0, iso)) < 0.f;
    if ((int)c0 + (int)c1 + (int)c2 != 2) return false;
    if (p == nullptr) return true;
    int n = 0;
    e0 = e1 = 0;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const bool c = e == 0 ? c0 : e == 1 ? c1 : c2;
        if (!c) continue;
        const int a = e, b = e == 2 ? 0 : e + 1;
        const float fa = fd.f[a], fb = fd.f[b];
        const float* va = V + 3ll * fd.v[a];
        const float* vb = V + 3ll * fd.v[b];
#pragma unroll
        for (int d = 0; d < 3; ++d) p[3 * n + d] = interp_iso(va[d], vb[d], fa, fb, iso);
        if (n == 0) e0 = e; else e1 = e;
        ++n;
    }
    return true;
}


// ---------------------------------------------------------------------------------------------------------
// The segment pool is written by `emit`, then read by `link` and `walk` WHILE the zero fill streams 12 K F bytes
// through the L2.  Pool accesses carry an L2 evict_last policy (the fill's stores are evict-first, __stcs), so the
// pool stays cache resident and the latency-bound chain does not queue behind the fill in DRAM.
__device__ __forceinline__ unsigned long long keep_policy() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ int ld_keep(const int* a, unsigned long long pol) {
    int v;
    asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_keep(const float* a, unsigned long long pol) {
    float v;
    asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_keep(int* a, int v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(a), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_keep(unsigned long long* a, unsigned long long v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(a), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_keep(float* a, float v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(a), "f"(v), "l"(pol) : "memory");
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// ---------------------------------------------------------------------------------------------------------
// prep: rank-sort the isovalues (ties keep input order), zero the per-call counters.
__global__ void contour_prep_kernel(const float* __restrict__ iso, int K, ContourWs w) {
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const float vi = iso[i];
        int rank = 0;
        for (int j = 0; j < K; ++j) {
            const float vj = iso[j];
            rank += (vj < vi) || (vj == vi && j < i);
        }
        // NaN isovalues never cross anything; park them (stably) at the end
        if (vi != vi) {
            int nn = 0, before = 0;
            for (int j = 0; j < K; ++j) { const float vj = iso[j]; nn += (vj == vj); before += (vj != vj) && j < i; }
            rank = nn + before;
        }
        w.iso_sorted[rank] = vi;
        w.iso_perm[rank] = i;
        w.seg_count[i] = 0;
        w.cursor[i] = 0;
        w.rows_used[i] = 0;
    }
    if (threadIdx.x < 16) w.hdr[threadIdx.x] = 0;
}

// count / emit: one thread per face.
template <bool EMIT>
__global__ void __launch_bounds__(kFaceThreads)
contour_faces_kernel(const float* __restrict__ V, const int* __restrict__ F, const float* __restrict__ fields,
                     int nV, int nF, int K, ContourWs w) {
    extern __shared__ float s_iso[];  // [K] sorted isovalues, then [K] perm as int
    int* s_perm = reinterpret_cast<int*>(s_iso + K);
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        s_iso[i] = w.iso_sorted[i];
        s_perm[i] = w.iso_perm[i];
    }
    __syncthreads();
    if (EMIT && w.hdr[0] != 0) return;  // pool overflow detected by the scan: nothing may be written

    const int lane = threadIdx.x & 31;
    const int hmask = EMIT ? (w.hdr[2] - 1) : 0;
    const unsigned long long keep = keep_policy();
    const int face_base = blockIdx.x * blockDim.x;
    // every warp runs whole: inactive lanes carry an empty isovalue window
    const int t = face_base + threadIdx.x;
    FaceData fd;
    int lo = 0, hi = 0;
    if (t < nF) {
        fd.v[0] = F[3ll * t]; fd.v[1] = F[3ll * t + 1]; fd.v[2] = F[3ll * t + 2];
        fd.f[0] = fields[fd.v[0]]; fd.f[1] = fields[fd.v[1]]; fd.f[2] = fields[fd.v[2]];
        const float mn = fminf(fd.f[0], fminf(fd.f[1], fd.f[2]));
        const float mx = fmaxf(fd.f[0], fmaxf(fd.f[1], fd.f[2]));
        // window = sorted positions with mn < iso < mx  (isocontours.cu:46-52 rejects everything else)
        int a = 0, b = K;
        while (a < b) { int m = (a + b) >> 1; if (s_iso[m] <= mn) a = m + 1; else b = m; }
        lo = a;
        b = K;
        while (a < b) { int m = (a + b) >> 1; if (s_iso[m] < mx) a = m + 1; else b = m; }
        hi = a;
    }
    for (int j = lo; __any_sync(0xffffffffu, j < hi); ++j) {
        bool valid = false;
        int k = -1;
        float p[6];
        int e0 = 0, e1 = 0;
        if (j < hi) {
            k = s_perm[j];
            valid = face_segment(fd, s_iso[j], V, EMIT ? p : nullptr, e0, e1);
        }
        // warp-aggregated slotting: one atomic per distinct isovalue present in the warp
        const int key = valid ? k : -1 - lane;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (!valid) continue;
        const int leader = __ffs(peers) - 1;
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (!EMIT) {
            if (lane == leader) atomicAdd(&w.seg_count[k], __popc(peers));
        } else {
            int basei = 0;
            if (lane == leader) basei = atomicAdd(&w.cursor[k], __popc(peers));
            basei = __shfl_sync(peers, basei, leader);
            const int slot = w.seg_offset[k] + basei + rank;
            st_keep(&w.seg_face[slot], t, keep);
            float* sp = w.seg_pts + 6ll * slot;
#pragma unroll
            for (int d = 0; d < 6; ++d) st_keep(sp + d, p[d], keep);
            w.used[slot] = 0;
            // register both endpoints under their (isovalue, mesh edge) key
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const int e = side == 0 ? e0 : e1;
                const int va = fd.v[e], vb = fd.v[e == 2 ? 0 : e + 1];
                const unsigned long long vmin = (unsigned)min(va, vb), vmax = (unsigned)max(va, vb);
                const unsigned long long hk = ((unsigned long long)k * (unsigned long long)nV + vmin) *
                                              (unsigned long long)nV + vmax;
                int hs = (int)(mix64(hk) & (unsigned long long)hmask);
                const int id = 2 * slot + side;
                while (true) {
                    const unsigned long long prev = atomicCAS(&w.hkeys[hs], kEmptyKey, hk);
                    if (prev == kEmptyKey || prev == hk) {
                        if (atomicCAS(&w.hvals[2 * hs], -1, id) != -1) atomicCAS(&w.hvals[2 * hs + 1], -1, id);
                        break;
                    }
                    hs = (hs + 1) & hmask;
                }
                st_keep(&w.seg_slot[id], hs, keep);
            }
        }
    }
}

// scan: exclusive prefix of the per-isovalue counts; sizes the hash table; flags pool overflow.
__global__ void contour_scan_kernel(int K, ContourWs w) {
    __shared__ int s_part[1024];
    __shared__ int s_total;
    const int tid = threadIdx.x;
    const int per = (K + blockDim.x - 1) / blockDim.x;
    const int b = tid * per, e = min(K, b + per);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += w.seg_count[i];
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int i = 0; i < (int)blockDim.x; ++i) { int v = s_part[i]; s_part[i] = run; run += v; }
        s_total = run;
    }
    __syncthreads();
    int run = s_part[tid];
    for (int i = b; i < e; ++i) { w.seg_offset[i] = run; run += w.seg_count[i]; }
    if (tid == 0) {
        const int total = s_total;
        w.seg_offset[K] = total;
        w.hdr[1] = total;
        w.hdr[3] = w.cap;
        w.hdr[0] = total > w.cap ? 1 : 0;
        long long want = 2ll * total;   // one key per crossed mesh edge (~ total): load factor 0.25 .. 0.5
        long long p = 1024;
        while (p < want) p <<= 1;
        if (p > w.hcap) p = w.hcap;
        w.hdr[2] = (int)p;
    }
}

__global__ void contour_hash_clear_kernel(ContourWs w) {
    const int n = w.hdr[2];
    const unsigned long long keep = keep_policy();   // the table is probed at random by `emit` and `link`: keep it in L2
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        st_keep(&w.hkeys[i], kEmptyKey, keep);
        st_keep(&w.hvals[2 * i], -1, keep);
        st_keep(&w.hvals[2 * i + 1], -1, keep);
    }
}

// link: partner endpoint through the shared mesh edge, kept only inside the reference's merge radius.
__global__ void contour_link_kernel(ContourWs w) {
    if (w.hdr[0] != 0) return;
    const int n = 2 * w.hdr[1];
    const unsigned long long keep = keep_policy();
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
        const int hs = ld_keep(&w.seg_slot[id], keep);
        const int a = ld_keep(&w.hvals[2 * hs], keep), b = ld_keep(&w.hvals[2 * hs + 1], keep);
        int partner = (a == id) ? b : (b == id ? a : -1);
        if (partner >= 0) {
            const float* p = w.seg_pts + 3ll * id;  // [seg][side][3] == flat [id][3]
            const float* q = w.seg_pts + 3ll * partner;
            const float dx = ld_keep(p, keep) - ld_keep(q, keep), dy = ld_keep(p + 1, keep) - ld_keep(q + 1, keep),
                        dz = ld_keep(p + 2, keep) - ld_keep(q + 2, keep);
            const float d = dx * dx + dy * dy + dz * dz;
            const float thr = 1e-6 * 1e-6;  // isocontours.cu:111
            if (!(d < thr)) partner = -1;
        }
        st_keep(&w.seg_link[id], partner, keep);
    }
}

// walk: one CTA per isovalue.  The isovalue's partner links are copied into shared memory (local endpoint ids) and
// one thread walks the chain with ONE dependent shared-memory load per step.  No `used` lookup sits on the chain:
// a walk can only ever run into a consumed segment through the back side (p1) of a SEED -- a non-seed segment is
// entered from the segment before it and left into the segment after it, both consumed by the same walk -- so
// seeding kills the one link that leads into the seed's p1 (link[partner(seed.p1)] = -1).  A closed loop coming
// back to its seed, or a later walk reaching an earlier seed from behind, then reads -1 and stops: exactly the
// reference's "no unused segment within the merge radius" exit (isocontours.cu:157-203).
constexpr int kWalkThreads = 512;
// Segments per isovalue handled entirely in shared memory (20 bytes per segment: 100 KB, two walk CTAs per SM, so
// K <= 296 isovalues are one wave).  Larger isovalues walk in the workspace.
constexpr int kWalkSmemSegs = 5120;
constexpr int kWalkSmemBytes = kWalkSmemSegs * 20;
// Second tier for multi-million-face meshes (F = 4 M, K = 200: up to ~10 k segments per isovalue): 210 KB, one CTA per SM.
constexpr int kWalkBigSegs = 11264;
constexpr int kWalkBigBytes = kWalkBigSegs * 20;
// list capacities per isovalue: contour pieces (closed loops + pieces of open chains) and open chains.  The small tier
// keeps its lists in shared memory; the big tier keeps them in the workspace (L1/L2 resident, touched by a few threads):
// at F = 4 M a few isovalues run along mesh edges and break into hundreds of pieces (the reference's 1e-6 merge radius),
// and their serial fallback (~2 ms each) would otherwise be the critical path of the whole call.
template <int CAP> struct RankCaps {
    static constexpr bool kGlobal = CAP > 5120;
    static constexpr int kLoops = kGlobal ? kBigMaxLoops : 320;
    static constexpr int kPaths = kGlobal ? kBigMaxPaths : 32;
};

// Where isovalue k's row stream goes, and how many rows fit: the dense [K, F, 3] contract of the reference
// (isocontours.cu:230-236), or the compact layout (2 S_k rows at 2 * seg_offset[k]; rows used = S_k + C_k <= 2 S_k).
__device__ __forceinline__ float* row_window(const ContourWs& w, int k, int off, int S, int nF, float* merged, int& limit) {
    if (w.compact) { limit = 2 * S; return merged + 6ll * off; }
    limit = nF;
    return merged + (size_t)k * nF * 3;
}

template <typename IdxT> struct WalkIdx;
template <> struct WalkIdx<int> {
    static constexpr int kNone = -1, kSep = -2;
    __device__ static bool none(int v) { return v < 0; }
    __device__ static bool is_point(int v) { return v >= 0; }
};
template <> struct WalkIdx<unsigned short> {
    static constexpr unsigned short kNone = 0xFFFFu, kSep = 0xFFFEu;
    __device__ static bool none(unsigned short v) { return v == kNone; }
    __device__ static bool is_point(unsigned short v) { return v < kSep; }
};

// Serial walk (isovalues with open chains, or too large for shared memory).  nx[e] = entry endpoint of the segment
// that follows the segment entered through endpoint e (link[e ^ 1]), so the chain is ONE dependent load per step;
// face[] is a private copy of the face ids so that the seed searches stay out of global memory when the working set
// is in shared memory (the zero fill saturates HBM while this runs).
template <typename IdxT, typename UsedT>
__device__ __forceinline__ void walk_isovalue(int S, int off, int nF, int k, const ContourWs& w, IdxT* nx, IdxT* order,
                                              UsedT* used, int* face_copy, const unsigned short* by_face,
                                              float* __restrict__ merged, int* __restrict__ nums) {
    using X = WalkIdx<IdxT>;
    __shared__ int s_face[kWalkThreads / 32];
    __shared__ int s_best[kWalkThreads / 32];
    __shared__ int s_seed;
    __shared__ int s_rows;
    const unsigned long long keep = keep_policy();
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const int a = ld_keep(&w.seg_link[2ll * off + 2 * i], keep), b = ld_keep(&w.seg_link[2ll * off + 2 * i + 1], keep);
        nx[2 * i] = b < 0 ? X::kNone : (IdxT)(b - 2 * off);       // all partners belong to the same isovalue
        nx[2 * i + 1] = a < 0 ? X::kNone : (IdxT)(a - 2 * off);
        used[i] = 0;
        if (face_copy) face_copy[i] = ld_keep(&w.seg_face[off + i], keep);
    }
    const int* face = face_copy ? face_copy : w.seg_face + off;
    const float* pts = w.seg_pts + 6ll * off;       // [S][2][3] == flat [endpoint][3]; rows <= S + C <= 2S
    int limit;
    float* out = row_window(w, k, off, S, nF, merged, limit);
    int row = 0, contours = 0;
    __syncthreads();
    if (by_face != nullptr) {
        // Seeds in ascending face order = one pass over the segments sorted by face id (sorted once, in parallel,
        // by the caller): the whole walk is a single thread touching shared memory only, no block-wide
        // synchronisation per contour piece.
        if (threadIdx.x == 0) {
            for (int c = 0; c < S; ++c) {
                const int seed = by_face[c];
                if (used[seed]) continue;
                if (contours > 0) order[row++] = X::kSep;  // separator between loops (isocontours.cu:135-140)
                order[row++] = (IdxT)(2 * seed);           // seed.p1 (isocontours.cu:143-146)
                used[seed] = 1;
                ++contours;
                const IdxT z = nx[2 * seed + 1];           // partner of seed.p1: the link into the seed dies
                if (!X::none(z)) nx[(int)z ^ 1] = X::kNone;
                int e = 2 * seed;                          // entered through p1, tail = p2 (isocontours.cu:151-153)
                while (true) {
                    const IdxT y = nx[e];                  // the only dependent load of the chain
                    if (X::none(y)) break;
                    e = (int)y;
                    order[row++] = (IdxT)(e ^ 1);          // left through its other endpoint: the emitted row
                    used[e >> 1] = 1;
                }
            }
        }
    } else {
    // phase 1: one thread records the row order (endpoint ids); nothing but shared-memory loads on its chain
    while (true) {
        // lowest face id among unused segments (ascending-face seed order); face ids are unique per isovalue
        int best_face = 0x7fffffff, best_seg = -1;
        for (int i = threadIdx.x; i < S; i += blockDim.x) {
            if (!used[i]) {
                const int f = face[i];
                if (f < best_face) { best_face = f; best_seg = i; }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const int of = __shfl_xor_sync(0xffffffffu, best_face, o);
            const int os = __shfl_xor_sync(0xffffffffu, best_seg, o);
            if (of < best_face) { best_face = of; best_seg = os; }
        }
        if ((threadIdx.x & 31) == 0) { s_face[threadIdx.x >> 5] = best_face; s_best[threadIdx.x >> 5] = best_seg; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int bf = s_face[0], bs = s_best[0];
            for (int i = 1; i < kWalkThreads / 32; ++i) if (s_face[i] < bf) { bf = s_face[i]; bs = s_best[i]; }
            s_seed = bs;
        }
        __syncthreads();
        const int seed = s_seed;
        if (seed < 0) break;
        if (threadIdx.x == 0) {
            if (contours > 0) order[row++] = X::kSep;  // separator between loops (isocontours.cu:135-140)
            order[row++] = (IdxT)(2 * seed);           // seed.p1 (isocontours.cu:143-146)
            used[seed] = 1;
            ++contours;
            // the link into the seed's p1 becomes a dead end: the loop closes when the walk comes back to it, and a
            // later walk that reaches this seed from behind stops there
            const IdxT z = nx[2 * seed + 1];           // partner of seed.p1
            if (!X::none(z)) nx[(int)z ^ 1] = X::kNone;
            int e = 2 * seed;                          // entered through p1, tail = p2 (isocontours.cu:151-153)
            while (true) {
                const IdxT y = nx[e];                  // entry endpoint of the next segment (the only dependent load)
                if (X::none(y)) break;
                e = (int)y;
                order[row++] = (IdxT)(e ^ 1);          // it is left through its other endpoint: the emitted row
                used[e >> 1] = 1;
            }
        }
        __syncthreads();  // `used` written by thread 0 is visible to the next seed search
    }
    }
    if (threadIdx.x == 0) {
        if (contours > 0) order[row++] = X::kSep;      // trailing separator (isocontours.cu:205-210)
        nums[k] = contours;
        s_rows = row;
        w.rows_used[k] = min(row, limit);
    }
    __syncthreads();
    // phase 2: all threads gather the points and write the row stream (or hand the row order to the scatter kernel)
    const int rows = min(s_rows, limit);
    if (w.defer) {
        int* gorder = w.seg_order + 2ll * off;
        if ((const void*)gorder != (const void*)order)
            for (int r = threadIdx.x; r < rows; r += blockDim.x) { const IdxT e = order[r]; gorder[r] = X::is_point(e) ? (int)e : -2; }
        return;
    }
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        const IdxT e = order[r];
        float a = -1e9f, b = -1e9f, c = -1e9f;
        if (X::is_point(e)) { const float* q = pts + 3ll * (int)e; a = ld_keep(q, keep); b = ld_keep(q + 1, keep); c = ld_keep(q + 2, keep); }
        out[3ll * r] = a; out[3ll * r + 1] = b; out[3ll * r + 2] = c;
    }
}

// Parallel walk of one isovalue.  The serial walk of the reference is a traversal of the functional graph
//   next(e) = link[e ^ 1]        on "states" e = endpoint through which a segment is entered,
// so its result can be computed with pointer doubling instead of one dependent load per segment:
//
// CLOSED loops (every endpoint has its partner -- watertight meshes).  Every undirected loop is two directed cycles of
// states; the reference starts at the loop's lowest face (ascending-face seed order) and enters that segment through
// p1, i.e. it follows the cycle that contains the EVEN state of the lowest-face segment.  One doubling pass carries,
// per state, (minimum over the next 2^t states of key = 2 * face + parity(state), distance to that minimum): after
// ceil(log2 S) rounds every state knows its loop's seed face, whether its cycle is the walked direction (key even) and
// how far behind the seed it is.
//
// OPEN chains (a partner is missing: mesh boundary, or a pair outside the reference's merge radius).  A first doubling
// pass gives every state its distance to the dead end of its directed path and the identity of that dead end; each
// chain is then laid out by position and the reference's seed order is replayed on INTERVALS: the lowest unused face
// of the chain seeds a piece that runs from there to the end of the remaining interval in the direction of the seed's
// p1 -> p2 traversal (the walk only stops at a dead end or at the back of an earlier seed, isocontours.cu:157-203), the
// rest of the interval is seeded again.  One block-wide arg-min per piece instead of one dependent load per segment.
//
// Pieces of both kinds are ordered by seed face, row offsets are a prefix sum of (length + 1), and every state of a walked
// direction writes its own output row: rank 0 -> the seed's p1, rank r >= 1 -> the exit endpoint of the r-th segment
// (isocontours.cu:143-203), then the (-1e9) separator row (:135-140, :205-210).
// Returns false (nothing written) when a list capacity is exceeded: the caller then walks the isovalue serially.
template <int CAP>
__device__ bool walk_ranked(int S, int off, int nF, int k, const ContourWs& w, unsigned char* smem,
                            float* __restrict__ merged, int* __restrict__ nums) {
    unsigned short* J = reinterpret_cast<unsigned short*>(smem);                 // [2S] jump pointers
    unsigned short* D = J + 2 * CAP;                                   // [2S] distances
    int* V = reinterpret_cast<int*>(D + 2 * CAP);                      // [2S] keys (closed) / dead-end info (open)
    unsigned short* NX = reinterpret_cast<unsigned short*>(V + 2 * CAP);  // [2S] next(e); later P | Q
    unsigned short* P = NX;                                                      // [S] state at a chain position
    unsigned short* Q = NX + CAP;                                      // [S] piece of a chain position
    constexpr unsigned short kNone = 0xFFFFu;
    constexpr int kMaxLoops = RankCaps<CAP>::kLoops, kMaxPaths = RankCaps<CAP>::kPaths;
    constexpr int kRPT = 2 * CAP / kWalkThreads;   // traversal states per thread (20 / 42)
    static_assert(2 * CAP < 32768, "state ids are packed into 15 bits of V");
    __shared__ int s_nloops, s_npaths, s_slot;
    constexpr bool kGlobalLists = RankCaps<CAP>::kGlobal;
    __shared__ int s_lists[kGlobalLists ? 1 : 7 * kMaxLoops + 3 * kMaxPaths];
    int* const lists = kGlobalLists ? w.walk_lists + (size_t)k * kWalkListInts : s_lists;
    int* const s_lface = lists;                      // piece: seed face (discovery order)
    int* const s_llen = lists + kMaxLoops;           //        length
    int* const s_linfo = lists + 2 * kMaxLoops;      // open piece: seed position << 1 | reverse flag
    int* const s_lbase = lists + 3 * kMaxLoops;      // piece: first output row
    int* const s_sface = lists + 4 * kMaxLoops;      // sorted by seed face
    int* const s_slen = lists + 5 * kMaxLoops;
    int* const s_sbase = lists + 6 * kMaxLoops;
    int* const s_pterm = lists + 7 * kMaxLoops;      // chain: dead end
    int* const s_plen = s_pterm + kMaxPaths;         //        length
    int* const s_pbase = s_pterm + 2 * kMaxPaths;    //        P offset
    __shared__ unsigned long long s_cmin[CAP / 32];   // per 32 chain positions: min (face << 32 | position)
    const int tid = threadIdx.x;
    const int n = 2 * S;
    const unsigned long long keep = keep_policy();
    int open_links = 0;
    for (int i = tid; i < S; i += kWalkThreads) {
        const int a = ld_keep(&w.seg_link[2ll * off + 2 * i], keep), b = ld_keep(&w.seg_link[2ll * off + 2 * i + 1], keep);
        open_links |= (a < 0) | (b < 0);
        const int f = ld_keep(&w.seg_face[off + i], keep);
        NX[2 * i] = b < 0 ? kNone : (unsigned short)(b - 2 * off);        // next(2i)   = link[2i + 1]
        NX[2 * i + 1] = a < 0 ? kNone : (unsigned short)(a - 2 * off);    // next(2i+1) = link[2i]
        V[2 * i] = 2 * f;
        V[2 * i + 1] = 2 * f + 1;
    }
    if (tid == 0) { s_nloops = 0; s_npaths = 0; }
    const bool mixed = __syncthreads_or(open_links) != 0;

    // in-place doubling round: read phase -> barrier -> write phase -> barrier.  MODE 0: distance to the dead end
    // (dead ends are self loops with distance 0); MODE 1: cycle minimum of V and distance to it (open states skipped).
    auto doubling = [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        for (int step = 1; step < S; step <<= 1) {
            unsigned int njd[kRPT];   // new jump pointer (low half) and distance (high half)
            int nv[kRPT];
#pragma unroll
            for (int q = 0; q < kRPT; ++q) {
                if (q * kWalkThreads >= n) break;   // uniform
                const int e = tid + q * kWalkThreads;
                if (e < n) {
                    if (MODE == 0) {
                        const int j = J[e];
                        njd[q] = (unsigned int)J[j] | ((unsigned int)(D[e] + D[j]) << 16);
                    } else {
                        const int ve = V[e];
                        if (ve < 0) { njd[q] = 0u; nv[q] = ve; }   // open state: untouched (its J / D hold a face id)
                        else {
                            const int j = J[e];
                            const unsigned int jj = J[j];
                            const int vj = V[j];
                            if (vj < ve) { nv[q] = vj; njd[q] = jj | ((unsigned int)(D[j] + step) << 16); }
                            else { nv[q] = ve; njd[q] = jj | ((unsigned int)D[e] << 16); }
                        }
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < kRPT; ++q) {
                if (q * kWalkThreads >= n) break;   // uniform
                const int e = tid + q * kWalkThreads;
                if (e < n) {
                    if (MODE == 0) { J[e] = (unsigned short)(njd[q] & 0xffffu); D[e] = (unsigned short)(njd[q] >> 16); }
                    else if (nv[q] >= 0) { J[e] = (unsigned short)(njd[q] & 0xffffu); D[e] = (unsigned short)(njd[q] >> 16); V[e] = nv[q]; }
                }
            }
            __syncthreads();
        }
    };

    if (mixed) {
        // ---- pass 1: dead end and distance to it for every state of an open chain --------------------------------------
        for (int e = tid; e < n; e += kWalkThreads) {
            const unsigned short nx = NX[e];
            J[e] = nx == kNone ? (unsigned short)e : nx;
            D[e] = nx == kNone ? 0 : 1;
        }
        __syncthreads();
        doubling(std::integral_constant<int, 0>{});
        // open state <=> its jump target is a dead end.  Keep (dead end, distance) in V (sign bit set) and the face id
        // in the now free (J, D) pair of the state.
        int vnew[kRPT];
#pragma unroll
        for (int q = 0; q < kRPT; ++q) {
            if (q * kWalkThreads >= n) break;
            const int e = tid + q * kWalkThreads;
            if (e < n) {
                const int t = J[e];
                vnew[q] = NX[t] == kNone ? (int)(0x80000000u | ((unsigned int)t << 16) | (unsigned int)D[e]) : V[e];
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kRPT; ++q) {
            if (q * kWalkThreads >= n) break;
            const int e = tid + q * kWalkThreads;
            if (e < n) {
                const int f = V[e] >> 1;          // V still holds 2 * face + parity
                if (vnew[q] < 0) { J[e] = (unsigned short)(f & 0xffff); D[e] = (unsigned short)((unsigned int)f >> 16); }
                V[e] = vnew[q];
            }
        }
        __syncthreads();
    }
    // ---- pass 2: closed loops ----------------------------------------------------------------------------------------------
    for (int e = tid; e < n; e += kWalkThreads) {
        if (V[e] >= 0) { J[e] = NX[e]; D[e] = 0; }
    }
    __syncthreads();
    doubling(std::integral_constant<int, 1>{});
    // seed states of closed loops: even states that are their own cycle minimum
    for (int i = tid; i < S; i += kWalkThreads) {
        const int e = 2 * i;
        if (V[e] >= 0 && D[e] == 0 && (V[e] & 1) == 0) {
            const int slot = atomicAdd(&s_nloops, 1);
            if (slot < kMaxLoops) {
                s_lface[slot] = V[e] >> 1;
                s_llen[slot] = (int)D[NX[e]] + 1;    // the state after the seed is length - 1 steps behind it
                s_linfo[slot] = -1;
            }
        }
    }
    __syncthreads();
    if (s_nloops > kMaxLoops) return false;

    if (mixed) {
        // ---- open chains: canonical direction = the one whose dead end has the smaller state id ------------------------
        auto dead_end = [&](int e) { return (int)(((unsigned int)V[e] >> 16) & 0x7fffu); };
        auto dist_end = [&](int e) { return (int)((unsigned int)V[e] & 0xffffu); };
        for (int e = tid; e < n; e += kWalkThreads) {
            if (V[e] < 0 && NX[e] == kNone && e < dead_end(e ^ 1)) {   // e is the dead end of a canonical direction
                const int slot = atomicAdd(&s_npaths, 1);
                if (slot < kMaxPaths) { s_pterm[slot] = e; s_plen[slot] = dist_end(e ^ 1) + 1; }
            }
        }
        __syncthreads();
        const int npaths = s_npaths;
        if (npaths > kMaxPaths) return false;
        if (tid == 0) { int run = 0; for (int c = 0; c < npaths; ++c) { s_pbase[c] = run; run += s_plen[c]; } }
        __syncthreads();   // NX is dead from here on: its storage becomes P | Q
        for (int e = tid; e < n; e += kWalkThreads) {
            if (V[e] >= 0) continue;
            const int t = dead_end(e);
            if (!(t < dead_end(e ^ 1))) continue;                  // the reverse traversal of its chain
            int c = 0;
            while (s_pterm[c] != t) ++c;
            P[s_pbase[c] + dist_end(e ^ 1)] = (unsigned short)e;   // position = distance of the reverse state to ITS dead end
        }
        __syncthreads();
        // faces by chain position, compact (the J array is free: closed loops are done with it, and the face ids that
        // the open states keep in their (J, D) pairs are read into registers first)
        int* facepos = reinterpret_cast<int*>(J);
        {
            const int total = s_pbase[npaths - 1] + s_plen[npaths - 1];
            int fr[(CAP + kWalkThreads - 1) / kWalkThreads];
#pragma unroll
            for (int q = 0; q < (CAP + kWalkThreads - 1) / kWalkThreads; ++q) {
                const int g = tid + q * kWalkThreads;
                if (g < total) { const int e = P[g]; fr[q] = (int)((unsigned int)J[e] | ((unsigned int)D[e] << 16)); }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < (CAP + kWalkThreads - 1) / kWalkThreads; ++q) {
                const int g = tid + q * kWalkThreads;
                if (g < total) facepos[g] = fr[q];
            }
            __syncthreads();
            // minima of aligned 32-position chunks: inside the unused interval of a chain they never change, so the
            // arg-min of a piece reads two partial chunks and a few chunk minima instead of the whole interval
            for (int c = tid >> 5; 32 * c < total; c += kWalkThreads / 32) {
                const int g = 32 * c + (tid & 31);
                unsigned long long key = g < total ? (((unsigned long long)(unsigned int)facepos[g] << 32) | (unsigned int)g) : ~0ull;
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                    key = other < key ? other : key;
                }
                if ((tid & 31) == 0) s_cmin[c] = key;
            }
        }
        // replay the seed order on intervals: one WARP per chain (chains are independent; no block-wide barrier per piece)
        if (tid == 0) s_slot = 0;   // overflow flag
        __syncthreads();

        const int lane = tid & 31;
        for (int c = tid >> 5; c < npaths; c += kWalkThreads / 32) {
            const int base = s_pbase[c];
            int lo = 0, hi = s_plen[c] - 1;
            while (lo <= hi) {
                // arg-min of the face id over the interval: two partial 32-position chunks + the chunk minima between
                unsigned int bf = 0xffffffffu;   // this lane's best face and its global position
                int bg = 0;
                {
                    const int glo = base + lo, ghi = base + hi;
                    const int c_lo = glo >> 5, c_hi = ghi >> 5;
                    int g = 32 * c_lo + lane;
                    if (g >= glo && g <= ghi) { bf = (unsigned int)facepos[g]; bg = g; }
                    if (c_hi > c_lo) {
                        g = 32 * c_hi + lane;
                        if (g <= ghi) { const unsigned int f = (unsigned int)facepos[g]; if (f < bf) { bf = f; bg = g; } }
                        for (int c2 = c_lo + 1 + lane; c2 < c_hi; c2 += 32) {
                            const unsigned long long key = s_cmin[c2];
                            const unsigned int f = (unsigned int)(key >> 32);
                            if (f < bf) { bf = f; bg = (int)(key & 0xffffffffu); }
                        }
                    }
                }
                const unsigned int fmin = __reduce_min_sync(0xffffffffu, bf);       // face ids are unique: one owner
                const int owner = __ffs(__ballot_sync(0xffffffffu, bf == fmin)) - 1;
                const int pmin = __shfl_sync(0xffffffffu, bg, owner) - base;
                const int e = P[base + pmin];
                const bool reverse = (e & 1) != 0;   // the seed's p1 -> p2 traversal runs against the canonical direction
                const int a0 = reverse ? lo : pmin, a1 = reverse ? pmin : hi;
                int slot = 0;
                if (lane == 0) slot = atomicAdd(&s_nloops, 1);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (slot >= kMaxLoops) { s_slot = 1; break; }
                if (lane == 0) {
                    s_lface[slot] = (int)fmin;
                    s_llen[slot] = a1 - a0 + 1;
                    s_linfo[slot] = ((base + pmin) << 1) | (reverse ? 1 : 0);
                }
                for (int pp = a0 + lane; pp <= a1; pp += 32) Q[base + pp] = (unsigned short)slot;
                if (reverse) lo = pmin + 1; else hi = pmin - 1;
            }
        }
        __syncthreads();
        if (s_slot) return false;
    }
    const int C = s_nloops;
    // order by seed face (counting rank; C is small), row bases by a serial prefix over the sorted pieces
    for (int c = tid; c < C; c += kWalkThreads) {
        const int f = s_lface[c];
        int r = 0;
        for (int o = 0; o < C; ++o) r += s_lface[o] < f;
        s_sface[r] = f;
        s_slen[r] = s_llen[c];
        s_lbase[c] = r;            // rank for now
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int c = 0; c < C; ++c) { s_sbase[c] = run; run += s_slen[c] + 1; }
        nums[k] = C;
        w.rows_used[k] = min(run, w.compact ? 2 * S : nF);
    }
    __syncthreads();
    for (int c = tid; c < C; c += kWalkThreads) s_lbase[c] = s_sbase[s_lbase[c]];
    __syncthreads();
    // ---- every state of a walked direction writes its row -------------------------------------------------------------------
    const float* pts = w.seg_pts + 6ll * off;
    int limit;
    float* out = row_window(w, k, off, S, nF, merged, limit);
    int* gorder = w.seg_order + 2ll * off;
    auto emit_row = [&](int row, int endpoint) {
        if (w.defer) { if (row < limit) gorder[row] = endpoint; return; }
        if (row < limit) {
            const float* q = pts + 3ll * endpoint;
            out[3ll * row] = ld_keep(q, keep); out[3ll * row + 1] = ld_keep(q + 1, keep); out[3ll * row + 2] = ld_keep(q + 2, keep);
        }
    };
    auto emit_separator = [&](int row) {
        if (w.defer) { if (row < limit) gorder[row] = -2; return; }
        if (row < limit) { float* sp = out + 3ll * row; sp[0] = -1e9f; sp[1] = -1e9f; sp[2] = -1e9f; }
    };
    for (int e = tid; e < n; e += kWalkThreads) {
        const int key = V[e];
        if (key < 0 || (key & 1)) continue;          // open chain (below) / the reverse traversal of its loop
        const int f = key >> 1;
        int lo = 0, hi = C - 1;                      // piece index: position of f among the sorted seed faces
        while (lo < hi) { const int m = (lo + hi) >> 1; if (s_sface[m] < f) lo = m + 1; else hi = m; }
        const int len = s_slen[lo], base = s_sbase[lo];
        const int d = D[e];
        const int r = d == 0 ? 0 : len - d;          // segments walked before this one
        emit_row(base + r, r == 0 ? e : (e ^ 1));
        if (r == 0) emit_separator(base + len);
    }
    if (mixed) {
        const int total = s_npaths > 0 ? s_pbase[s_npaths - 1] + s_plen[s_npaths - 1] : 0;
        for (int g = tid; g < total; g += kWalkThreads) {
            const int e = P[g];                      // canonical state of the segment at chain position g
            const int piece = Q[g];
            const int info = s_linfo[piece];
            const int seed_pos = info >> 1;
            const bool reverse = info & 1;
            const int r = reverse ? seed_pos - g : g - seed_pos;
            const int base = s_lbase[piece];
            // rank 0: the seed's p1 (its even endpoint); rank >= 1: the endpoint the segment is left through
            emit_row(base + r, r == 0 ? (e & ~1) : (reverse ? e : (e ^ 1)));
            if (r == 0) emit_separator(base + s_llen[piece]);
        }
    }
    return true;
}

// Two tiers of the shared-memory walk: CAP = kWalkSmemSegs (100 KB, two CTAs per SM) and CAP = kWalkBigSegs (210 KB, one
// CTA per SM) for the isovalues of multi-million-face meshes; a launch of tier T handles exactly the isovalues whose
// segment count falls into its range and leaves the others to the other launch.
template <int CAP, int MIN_CTAS, bool BIG>
__global__ void __launch_bounds__(kWalkThreads, MIN_CTAS)
contour_walk_kernel(int nF, int two_tiers, ContourWs w, float* __restrict__ merged, int* __restrict__ nums) {
    extern __shared__ int s_dyn[];
    const int k = blockIdx.x;
    const long long t_begin = clock64();
    struct Stamp { int* slot; long long t0; __device__ ~Stamp() { if (threadIdx.x == 0) *slot = (int)(clock64() - t0); } } stamp{&w.cursor[k], t_begin};
    const bool overflow = w.hdr[0] != 0;
    const int S = overflow ? 0 : w.seg_count[k];
    const int off = w.seg_offset[k];
    if (BIG ? (S <= kWalkSmemSegs || S > kWalkBigSegs) : (two_tiers && S > kWalkSmemSegs && S <= kWalkBigSegs)) {
        stamp.slot = &w.hdr[15];   // not this tier's isovalue: leave its time stamp alone
        return;
    }
    if (S == 0) {
        if (threadIdx.x == 0) { nums[k] = 0; w.rows_used[k] = 0; }
        return;
    }
    // Working set of the chain walk: partner links (local endpoint ids), the row order being recorded and the
    // per-segment consumed flags.  In shared memory when the isovalue fits (no global traffic on the chain, which
    // matters because the zero fill saturates HBM concurrently), in the workspace otherwise.
    if (S <= CAP) {
        if (walk_ranked<CAP>(S, off, nF, k, w, reinterpret_cast<unsigned char*>(s_dyn), merged, nums)) return;
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(&w.hdr[8], 1);   // statistics: isovalues walked serially
        unsigned short* nx = reinterpret_cast<unsigned short*>(s_dyn);
        unsigned short* order = nx + 2 * CAP;
        int* face = reinterpret_cast<int*>(order + 2 * CAP);
        unsigned char* used = reinterpret_cast<unsigned char*>(face + CAP);
        walk_isovalue<unsigned short, unsigned char>(S, off, nF, k, w, nx, order, used, face, nullptr, merged, nums);
    } else {
        // the hash-slot array is dead after `link`: reuse it as int flags
        if (threadIdx.x == 0) atomicAdd(&w.hdr[8], 1);
        walk_isovalue<int, int>(S, off, nF, k, w, w.seg_link + 2ll * off, w.seg_order + 2ll * off, w.seg_slot + 2ll * off,
                                nullptr, nullptr, merged, nums);
    }
}

// Both tiers of the walk.  A CTA of the big tier needs a whole SM (210 KB of shared memory, 58 k registers): while the
// zero fill of the dense contract occupies the SMs its CTAs -- even the ones that exit at once -- are only placed as the
// fill drains.  It is therefore launched only for meshes large enough to have isovalues beyond the small tier
// (S_k grows like sqrt(F): 4.6 k at F = 1 M, 9.4 k at F = 4 M on the closed bench mesh); below that threshold an isovalue
// with more than kWalkSmemSegs segments is walked serially, as before.
inline cudaError_t launch_walk(int nF, int K, const ContourWs& w, float* out, int* nums, cudaStream_t st) {
    auto small = contour_walk_kernel<kWalkSmemSegs, 2, false>;
    auto big = contour_walk_kernel<kWalkBigSegs, 1, true>;
    const bool use_big = nF >= kWalkBigMinFaces && w.walk_lists != nullptr;
    cudaError_t e = cudaFuncSetAttribute(small, cudaFuncAttributeMaxDynamicSharedMemorySize, kWalkSmemBytes);
    if (e != cudaSuccess) return e;
    small<<<K, kWalkThreads, kWalkSmemBytes, st>>>(nF, use_big ? 1 : 0, w, out, nums);
    if (use_big) {
        e = cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, kWalkBigBytes);
        if (e != cudaSuccess) return e;
        big<<<K, kWalkThreads, kWalkBigBytes, st>>>(nF, 1, w, out, nums);
    }
    return cudaGetLastError();
}

// Zero fill of the dense contract.  PHASE 0 (side stream, before the walk): rows [min(2 S_k, F), F).
// PHASE 1 (after the walk): rows [rows_used_k, min(2 S_k, F)).  On overflow everything is zeroed by phase 0.
template <int PHASE>
__global__ void __launch_bounds__(256)
contour_fill_kernel(int nF, int K, ContourWs w, float* __restrict__ merged) {
    const bool overflow = w.hdr[0] != 0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = blockIdx.y; k < K; k += gridDim.y) {
        const long long S = overflow ? 0 : w.seg_count[k];
        const long long split = min(2 * S, (long long)nF);
        long long r0, r1;
        if (PHASE == 0) { r0 = split; r1 = nF; } else { r0 = w.rows_used[k]; r1 = split; }
        if (r0 >= r1) continue;
        float* base = merged + (size_t)k * nF * 3;
        const long long e0 = 3 * r0, e1 = 3 * r1;  // float range inside this isovalue's slab
        // align the body to 16 bytes (the slab start is only 4-byte aligned in general)
        const uintptr_t addr0 = reinterpret_cast<uintptr_t>(base + e0);
        long long head = ((16 - (addr0 & 15)) & 15) >> 2;
        if (head > e1 - e0) head = e1 - e0;
        const long long body0 = e0 + head;
        const long long nvec = (e1 - body0) >> 2;
        const long long tail0 = body0 + 4 * nvec;
        if (tid < head) base[e0 + tid] = 0.f;
        if (tid < e1 - tail0) base[tail0 + tid] = 0.f;
        float4* vb = reinterpret_cast<float4*>(base + body0);
#pragma unroll 4
        for (long long i = tid; i < nvec; i += nth) __stcs(vb + i, z);
    }
}

static cudaStream_t g_side[64] = {nullptr};
static std::mutex g_mu;

// One non-blocking helper stream per device, created on first use (the zero fill forks onto it).
int get_side_stream(cudaStream_t* out) {
    int dev = 0;
    INF3DP_CUDA(cudaGetDevice(&dev));
    INF3DP_CHECK_ARG(dev >= 0 && dev < 64, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_side[dev] == nullptr) {
        // HIGHEST priority: the latency-bound chain (hash/link/walk) runs here while the bulk zero fill occupies the
        // caller's stream; the block scheduler only lets a later kernel overtake the pending blocks of an earlier one
        // when its stream has a higher priority (caller streams -- torch's default stream -- have the lowest)
        int least = 0, greatest = 0;
        INF3DP_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        INF3DP_CUDA(cudaStreamCreateWithPriority(&g_side[dev], cudaStreamNonBlocking, greatest));
    }
    *out = g_side[dev];
    return INF3DP_OK;
}

// Fork / join of the caller's stream and the helper stream.  The two events come from a small per-process pool (no
// cudaEventCreate / Destroy per call) and go back to it on EVERY exit path; if the call fails after the fork, the
// destructor still joins the helper stream into the caller's stream, so an error never leaves the two streams unordered.
static std::vector<cudaEvent_t> g_event_pool;

cudaEvent_t pool_get_event() {
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    }
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return e;
}

void pool_put_event(cudaEvent_t e) {
    if (!e) return;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_event_pool.size() < 64) g_event_pool.push_back(e); else cudaEventDestroy(e);
}

struct ForkJoin {
    cudaStream_t main_ = nullptr, side_ = nullptr;
    cudaEvent_t fork_ = nullptr, join_ = nullptr;
    bool forked = false, joined = false;
    ForkJoin(cudaStream_t m, cudaStream_t s) : main_(m), side_(s), fork_(pool_get_event()), join_(pool_get_event()) {}
    bool ok() const { return fork_ && join_; }
    cudaError_t fork() {
        cudaError_t e = cudaEventRecord(fork_, main_);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(side_, fork_, 0);
        forked = (e == cudaSuccess);
        return e;
    }
    cudaError_t join() {
        cudaError_t e = cudaEventRecord(join_, side_);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(main_, join_, 0);
        joined = true;
        return e;
    }
    ~ForkJoin() {
        if (forked && !joined) join();
        pool_put_event(fork_);
        pool_put_event(join_);
    }
};

// The face kernels keep the K sorted isovalues and their permutation in dynamic shared memory (8 K bytes): above the
// 48 KB default (K > 6144; the ABI accepts K <= 12000 = 96 KB) the kernel has to be opted in.
template <bool EMIT>
int faces_smem_opt_in(size_t fsmem);

// Largest pool that fits the caller's workspace (>= the default, doubling).
// Zero fill of the whole dense contract (12 K F bytes), independent of every other kernel of the call: it starts at t = 0
// on the caller's stream while the chain runs on the helper stream.  Many short blocks (see contour_fill_kernel).
__global__ void __launch_bounds__(256)
contour_fill_all_kernel(float* __restrict__ merged, long long nfloat) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    const uintptr_t addr0 = reinterpret_cast<uintptr_t>(merged);
    long long head = ((16 - (addr0 & 15)) & 15) >> 2;
    if (head > nfloat) head = nfloat;
    const long long nvec = (nfloat - head) >> 2;
    const long long tail0 = head + 4 * nvec;
    if (tid < head) merged[tid] = 0.f;
    if (tid < nfloat - tail0) merged[tail0 + tid] = 0.f;
    float4* vb = reinterpret_cast<float4*>(merged + head);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (long long i = tid; i < nvec; i += nth) __stcs(vb + i, z);
}

// Rows of the dense contract from the recorded row order (after the fill): out[k][r] = point of endpoint order[r], or the
// (-1e9) separator.
__global__ void __launch_bounds__(256)
contour_scatter_kernel(int nF, int K, ContourWs w, float* __restrict__ merged) {
    const unsigned long long keep = keep_policy();
    for (int k = blockIdx.y; k < K; k += gridDim.y) {
        const int rows = w.hdr[0] != 0 ? 0 : w.rows_used[k];
        const int off = w.seg_offset[k];
        const int* order = w.seg_order + 2ll * off;
        const float* pts = w.seg_pts + 6ll * off;
        float* out = merged + (size_t)k * nF * 3;
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
            const int e = ld_keep(order + r, keep);
            float a = -1e9f, b = -1e9f, c = -1e9f;
            if (e >= 0) { const float* q = pts + 3ll * e; a = ld_keep(q, keep); b = ld_keep(q + 1, keep); c = ld_keep(q + 2, keep); }
            out[3ll * r] = a; out[3ll * r + 1] = b; out[3ll * r + 2] = c;
        }
    }
}

// compact layout bookkeeping for the caller: row offset / row count of every isovalue
__global__ void contour_export_kernel(int K, ContourWs w, long long* __restrict__ row_offsets, int* __restrict__ row_counts) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) {
        row_offsets[k] = 2ll * w.seg_offset[k];
        row_counts[k] = w.hdr[0] != 0 ? 0 : w.rows_used[k];
    }
}

template <bool EMIT>
int faces_smem_opt_in(size_t fsmem) {
    if (fsmem > 48 * 1024)
        INF3DP_CUDA(cudaFuncSetAttribute(contour_faces_kernel<EMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
    return INF3DP_OK;
}

int fit_capacity(int nF, int K, size_t ws_bytes) {
    int cap = seg_capacity_default(nF, K);
    const long long worst = (long long)nF * (long long)K;
    while ((long long)cap * 2 <= worst && cap < 0x20000000 && carve(nullptr, K, cap * 2).total_bytes <= ws_bytes) cap *= 2;
    return cap;
}

}  // namespace
}  // namespace inf3dp

using namespace inf3dp;

extern "C" {

size_t inf3dp_extract_isocontours_workspace_bytes(int nV, int nF, int K) {
    (void)nV;
    if (nF < 0 || K < 0) return 0;
    const int k1 = K > 0 ? K : 1;
    return carve(nullptr, k1, seg_capacity_default(nF, k1)).total_bytes;
}

size_t inf3dp_extract_isocontours_workspace_bytes_segments(int nV, int nF, int K, int64_t min_segments) {
    (void)nV;
    if (nF < 0 || K < 0 || min_segments < 0) return 0;
    const int k1 = K > 0 ? K : 1;
    long long cap = seg_capacity_default(nF, k1);
    while (cap < min_segments && cap < 0x20000000) cap *= 2;
    return carve(nullptr, k1, (int)cap).total_bytes;
}

int inf3dp_extract_isocontours(const float* V, const int32_t* F, const float* fields, const float* iso, int nV,
                               int nF, int K, float* merged, int32_t* nums, void* ws, size_t ws_bytes,
                               inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(nV >= 0 && nF >= 0 && K >= 0, "extract_isocontours: negative size");
    if (K == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(iso && nums && ws, "extract_isocontours: null pointer");
    INF3DP_CHECK_ARG(nF == 0 || (V && F && fields && merged), "extract_isocontours: null pointer");
    INF3DP_CHECK_ARG(K <= 12000, "extract_isocontours: K=%d isovalues exceed the supported 12000", K);
    INF3DP_CHECK_ARG((double)K * (double)nV * (double)nV < 9.0e18, "extract_isocontours: K*nV^2 overflows the edge key");
    INF3DP_CHECK_ARG((reinterpret_cast<uintptr_t>(merged) & 3) == 0, "extract_isocontours: output not 4-byte aligned");
    ContourWs w = carve(ws, K, fit_capacity(nF, K, ws_bytes));
    if (ws_bytes < w.total_bytes) {
        set_error("extract_isocontours: workspace has %zu bytes, need %zu", ws_bytes, w.total_bytes);
        return INF3DP_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaStream_t side;
    int rc = get_side_stream(&side);
    if (rc != INF3DP_OK) return rc;
    // developer instrumentation (INF3DP_CONTOUR_DEBUG=1): per-phase CUDA-event timeline on both streams
    static const bool debug = getenv("INF3DP_CONTOUR_DEBUG") != nullptr;
    static const int dev_skip = getenv("INF3DP_CONTOUR_SKIP") ? atoi(getenv("INF3DP_CONTOUR_SKIP")) : 0;  // 1: no fill, 2: no chain
    cudaEvent_t dbg_ev[8] = {nullptr};
    if (debug) for (int i = 0; i < 8; ++i) cudaEventCreate(&dbg_ev[i]);
    auto mark = [&](int i, cudaStream_t s_) { if (debug) cudaEventRecord(dbg_ev[i], s_); };
    mark(0, st);

    const int sms = num_sms();
    const size_t fsmem = (size_t)K * 8;
    if ((rc = faces_smem_opt_in<false>(fsmem)) != INF3DP_OK) return rc;
    if ((rc = faces_smem_opt_in<true>(fsmem)) != INF3DP_OK) return rc;
    ForkJoin fj(st, side);     // joins and returns its events to the pool on every exit path
    if (!fj.ok()) { set_error("extract_isocontours: cudaEventCreate failed"); return INF3DP_ECUDA; }
    w.defer = 1;

    // fork at t = 0.  Helper stream (highest priority): the whole latency-bound chain prep -> count -> scan -> hash
    // clear -> emit -> link -> walk, which only RECORDS the row order of every isovalue.  Caller's stream: the zero fill
    // of the dense contract, which depends on nothing.  Join, then the recorded rows are scattered over the zeros.
    INF3DP_CUDA(fj.fork());
    mark(1, st);
    // (the fill is enqueued first: seven chain launches cost ~30 us of host time that it would otherwise wait for)
    mark(2, st);
    if (nF > 0 && dev_skip != 1) {
        // Many SHORT blocks (each thread streams ~16 float4 stores): SM slots free up continuously, so the
        // higher-priority chain kernels on the helper stream are dispatched between them.  (Measured alternative:
        // a persistent 2-CTA/SM fill cannot keep enough stores in flight -- 3.8 TB/s instead of 6.1 TB/s.)
        const long long nfloat = 3ll * (long long)nF * (long long)K;
        long long blocks = (nfloat / 4 + 4095) / 4096;
        if (blocks < 1) blocks = 1;
        if (blocks > 65535ll * 16) blocks = 65535ll * 16;
        contour_fill_all_kernel<<<(unsigned)blocks, 256, 0, st>>>(merged, nfloat);
        INF3DP_LAUNCH_CHECK("contour_fill_all_kernel");
    }
    const int fblocks = (nF + kFaceThreads - 1) / kFaceThreads;
    if (dev_skip != 2) {
        contour_prep_kernel<<<1, 1024, 0, side>>>(iso, K, w);
        INF3DP_LAUNCH_CHECK("contour_prep_kernel");
        if (fblocks > 0) {
            contour_faces_kernel<false><<<fblocks, kFaceThreads, fsmem, side>>>(V, F, fields, nV, nF, K, w);
            INF3DP_LAUNCH_CHECK("contour_faces_kernel<count>");
        }
        contour_scan_kernel<<<1, 1024, 0, side>>>(K, w);
        INF3DP_LAUNCH_CHECK("contour_scan_kernel");
        if (nF > 0) {
            contour_hash_clear_kernel<<<sms * 4, 256, 0, side>>>(w);
            INF3DP_LAUNCH_CHECK("contour_hash_clear_kernel");
            contour_faces_kernel<true><<<fblocks, kFaceThreads, fsmem, side>>>(V, F, fields, nV, nF, K, w);
            INF3DP_LAUNCH_CHECK("contour_faces_kernel<emit>");
            mark(4, side);
            contour_link_kernel<<<sms * 4, 256, 0, side>>>(w);
            INF3DP_LAUNCH_CHECK("contour_link_kernel");
            mark(5, side);
        }
        INF3DP_CUDA(launch_walk(nF, K, w, merged, nums, side));
        INF3DP_LAUNCH_CHECK("contour_walk_kernel");
    }
    mark(6, side);
    mark(3, st);
    INF3DP_CUDA(fj.join());
    if (nF > 0 && dev_skip != 2) {
        dim3 grid(8, K);
        contour_scatter_kernel<<<grid, 256, 0, st>>>(nF, K, w, merged);
        INF3DP_LAUNCH_CHECK("contour_scatter_kernel");
    }
    mark(7, st);
    if (debug) {
        cudaStreamSynchronize(st);
        float t[8] = {0};
        for (int i = 1; i < 8; ++i) cudaEventElapsedTime(&t[i], dbg_ev[0], dbg_ev[i]);
        int hdr_host[32] = {0};
        cudaMemcpy(hdr_host, w.hdr, sizeof(hdr_host), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[inf3dp contour debug] isovalues walked serially (list capacity exceeded or too large): %d of %d "
                        "(largest closed-loop count over capacity %d, open-chain count over capacity %d, piece overflow in replay %d)\n",
                hdr_host[8], K, hdr_host[9], hdr_host[10], hdr_host[11]);
        {
            std::vector<int> cyc(K), cnt(K);
            cudaMemcpy(cyc.data(), w.cursor, K * sizeof(int), cudaMemcpyDeviceToHost);
            cudaMemcpy(cnt.data(), w.seg_count, K * sizeof(int), cudaMemcpyDeviceToHost);
            int kmax = 0; long long sum = 0; int smax = 0;
            for (int i = 0; i < K; ++i) { if (cyc[i] > cyc[kmax]) kmax = i; sum += cyc[i]; if (cnt[i] > smax) smax = cnt[i]; }
            fprintf(stderr, "[inf3dp contour debug] walk CTA cycles: mean %.0f, max %d (isovalue %d, S=%d); max S over isovalues %d\n",
                    (double)sum / (K > 0 ? K : 1), cyc[kmax], kmax, cnt[kmax], smax);
        }
        fprintf(stderr, "[inf3dp contour debug] us since start: fork %.0f | caller stream: fill start %.0f end %.0f | chain stream: emit end %.0f "
                        "link end %.0f walk end %.0f | scatter end %.0f\n", t[1] * 1e3, t[2] * 1e3, t[3] * 1e3, t[4] * 1e3,
                t[5] * 1e3, t[6] * 1e3, t[7] * 1e3);
        for (int i = 0; i < 8; ++i) cudaEventDestroy(dbg_ev[i]);
    }
    return INF3DP_OK;
}

int64_t inf3dp_extract_isocontours_compact_rows(int nV, int nF, int K, size_t ws_bytes) {
    (void)nV;
    if (nF < 0 || K <= 0) return 0;
    return 2ll * (int64_t)fit_capacity(nF, K, ws_bytes);
}

int inf3dp_extract_isocontours_compact(const float* V, const int32_t* F, const float* fields, const float* iso, int nV,
                                       int nF, int K, float* rows, int64_t rows_capacity, int64_t* row_offsets,
                                       int32_t* row_counts, int32_t* nums, void* ws, size_t ws_bytes,
                                       inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(nV >= 0 && nF >= 0 && K >= 0, "extract_isocontours_compact: negative size");
    if (K == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(iso && nums && ws && row_offsets && row_counts, "extract_isocontours_compact: null pointer");
    INF3DP_CHECK_ARG(nF == 0 || (V && F && fields && rows), "extract_isocontours_compact: null pointer");
    INF3DP_CHECK_ARG(K <= 12000, "extract_isocontours_compact: K=%d isovalues exceed the supported 12000", K);
    INF3DP_CHECK_ARG((double)K * (double)nV * (double)nV < 9.0e18, "extract_isocontours_compact: K*nV^2 overflows the edge key");
    ContourWs w = carve(ws, K, fit_capacity(nF, K, ws_bytes));
    if (ws_bytes < w.total_bytes) {
        set_error("extract_isocontours_compact: workspace has %zu bytes, need %zu", ws_bytes, w.total_bytes);
        return INF3DP_EWORKSPACE;
    }
    INF3DP_CHECK_ARG(rows_capacity >= 2ll * w.cap, "extract_isocontours_compact: rows buffer holds %lld rows, need %lld "
                     "(inf3dp_extract_isocontours_compact_rows)", (long long)rows_capacity, 2ll * w.cap);
    w.compact = 1;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t fsmem = (size_t)K * 8;
    {
        int rc = faces_smem_opt_in<false>(fsmem);
        if (rc == INF3DP_OK) rc = faces_smem_opt_in<true>(fsmem);
        if (rc != INF3DP_OK) return rc;
    }
    contour_prep_kernel<<<1, 1024, 0, st>>>(iso, K, w);
    INF3DP_LAUNCH_CHECK("contour_prep_kernel");
    const int fblocks = (nF + kFaceThreads - 1) / kFaceThreads;
    if (fblocks > 0) {
        contour_faces_kernel<false><<<fblocks, kFaceThreads, fsmem, st>>>(V, F, fields, nV, nF, K, w);
        INF3DP_LAUNCH_CHECK("contour_faces_kernel<count>");
    }
    contour_scan_kernel<<<1, 1024, 0, st>>>(K, w);
    INF3DP_LAUNCH_CHECK("contour_scan_kernel");
    if (nF > 0) {
        const int sms = num_sms();
        contour_hash_clear_kernel<<<sms * 4, 256, 0, st>>>(w);
        INF3DP_LAUNCH_CHECK("contour_hash_clear_kernel");
        contour_faces_kernel<true><<<fblocks, kFaceThreads, fsmem, st>>>(V, F, fields, nV, nF, K, w);
        INF3DP_LAUNCH_CHECK("contour_faces_kernel<emit>");
        contour_link_kernel<<<sms * 4, 256, 0, st>>>(w);
        INF3DP_LAUNCH_CHECK("contour_link_kernel");
    }
    INF3DP_CUDA(launch_walk(nF, K, w, rows, nums, st));
    INF3DP_LAUNCH_CHECK("contour_walk_kernel");
    contour_export_kernel<<<(K + 255) / 256, 256, 0, st>>>(K, w, (long long*)row_offsets, row_counts);
    INF3DP_LAUNCH_CHECK("contour_export_kernel");
    return INF3DP_OK;
}

int64_t inf3dp_extract_isocontours_required_segments(const void* ws, int K, inf3dp_stream_t stream) {
    if (ws == nullptr || K <= 0) return 0;
    ContourWs w = carve(const_cast<void*>(ws), K, 1024);
    int hdr[4] = {0, 0, 0, 0};
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemcpyAsync(hdr, w.hdr, sizeof(hdr), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (int64_t)hdr[1];
}

int inf3dp_extract_isocontours_status(const void* ws, int K, int32_t* seg_counts_host, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(ws != nullptr && K >= 0, "extract_isocontours_status: bad argument");
    if (K == 0) return INF3DP_OK;
    // the pool capacity only affects offsets after the K-sized arrays, which is all this function reads
    ContourWs w = carve(const_cast<void*>(ws), K, 1024);
    int hdr[4];
    cudaStream_t st = (cudaStream_t)stream;
    INF3DP_CUDA(cudaMemcpyAsync(hdr, w.hdr, sizeof(hdr), cudaMemcpyDeviceToHost, st));
    if (seg_counts_host)
        INF3DP_CUDA(cudaMemcpyAsync(seg_counts_host, w.seg_count, (size_t)K * 4, cudaMemcpyDeviceToHost, st));
    INF3DP_CUDA(cudaStreamSynchronize(st));
    if (hdr[0] != 0) {
        set_error("extract_isocontours: %d segments exceed the workspace pool of %d; re-run with a workspace sized for "
                  "at least that many segments", hdr[1], hdr[3]);
        return INF3DP_EOVERFLOW;
    }
    return INF3DP_OK;
}

}  // extern "C"
