/* ORACLE (test infrastructure, not product code).
 *
 * Single-thread CPU restatement of the reference's two native contour operators, written from the
 * behaviour of /root/reference/contour_cuda (file:line cited per function).  Only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the shipped CUDA path never does.
 *
 * Pinning: the reference ships no golden vectors for these operators.  The
 * restatement is pinned against the reference extension itself, compiled unmodified-in-behaviour from
 * /root/reference/contour_cuda by oracle/build_ref.py into oracle/_ref/ and run on a B200
 * (tests/golden/make_contour_golden.py -> tests/golden/contour_ref_*.npz).
 *
 * Scheduling note: the reference's slot order inside one isovalue is decided by atomicAdd
 * (isocontours.cu:85) and is not deterministic.  This restatement uses ONE valid schedule: faces in
 * ascending index.  The intersection kernel's racing cell writes (intersection.cu:119-124,216-218) are
 * resolved by `winner_rule` (0: lowest linear voxel index wins, 1: highest wins), and
 * oracle_intersection_unmatched() checks an arbitrary race outcome against every candidate voxel.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off (no fast-math) so float expressions round exactly as
 * written; where nvcc's default -fmad=true contracts a*b+c in the reference build, fmaf() is written
 * explicitly and commented.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* isocontours.cu:9-11 -- v0 + (iso - f0) * (v1 - v0) / (f1 - f0); the product feeds a division, so no
 * mul+add contraction is possible and every operation is a correctly rounded IEEE single op. */
static inline float interp_iso(float v0, float v1, float f0, float f1, float iso) {
    return v0 + (iso - f0) * (v1 - v0) / (f1 - f0);
}

/* isocontours.cu:42-92 for one (face, isovalue): returns 1 and fills p[6] when exactly two edges cross. */
static int face_segment(const float* V, const int32_t* tri, const float* fields, float iso, float* p) {
    int v0 = tri[0], v1 = tri[1], v2 = tri[2];
    float f0 = fields[v0], f1 = fields[v1], f2 = fields[v2];
    int b0 = f0 <= iso, b1 = f1 <= iso, b2 = f2 <= iso;
    int a0 = f0 >= iso, a1 = f1 >= iso, a2 = f2 >= iso;
    if ((b0 && b1 && b2) || (a0 && a1 && a2)) return 0;                       /* :46-52 */
    float pts[9];
    int pc = 0;
    if ((f0 - iso) * (f1 - iso) < 0) {                                        /* :59 edge v0-v1 */
        for (int i = 0; i < 3; ++i) pts[pc++] = interp_iso(V[3 * v0 + i], V[3 * v1 + i], f0, f1, iso);
    }
    if ((f1 - iso) * (f2 - iso) < 0) {                                        /* :67 edge v1-v2 */
        for (int i = 0; i < 3; ++i) pts[pc++] = interp_iso(V[3 * v1 + i], V[3 * v2 + i], f1, f2, iso);
    }
    if ((f2 - iso) * (f0 - iso) < 0) {                                        /* :75 edge v2-v0 */
        for (int i = 0; i < 3; ++i) pts[pc++] = interp_iso(V[3 * v2 + i], V[3 * v0 + i], f2, f0, iso);
    }
    if (pc != 6) return 0;                                                    /* :83 */
    memcpy(p, pts, 6 * sizeof(float));
    return 1;
}

/* extract_isocontours (ext.cpp:5-17 -> isocontours.cu:216-293).
 * merged: [K, nF, 3] (zero-filled here, like torch::zeros at :234), nums: [K], seg_counts: [K] or NULL
 * (the reference's pairpoint_nums, not returned by its API but useful to tests). Returns 0, or -1 on OOM. */
int oracle_extract_isocontours(const float* V, const int32_t* F, const float* fields, const float* iso,
                               int nV, int nF, int K, float* merged, int32_t* nums, int32_t* seg_counts) {
    (void)nV;
    float* seg = (float*)malloc((size_t)(nF > 0 ? nF : 1) * 6 * sizeof(float));
    unsigned char* used = (unsigned char*)malloc((size_t)(nF > 0 ? nF : 1));
    if (!seg || !used) { free(seg); free(used); return -1; }
    memset(merged, 0, (size_t)K * nF * 3 * sizeof(float));
    const float thr = 1e-6 * 1e-6;                                            /* :111 (double product -> float) */
    for (int k = 0; k < K; ++k) {
        nums[k] = 0;
        int S = 0;
        for (int t = 0; t < nF; ++t)                                          /* one valid atomicAdd schedule */
            if (face_segment(V, F + 3 * t, fields, iso[k], seg + 6 * S)) ++S;
        if (seg_counts) seg_counts[k] = S;
        if (S == 0) continue;                                                 /* :116 */
        memset(used, 0, (size_t)S);
        float* out = merged + (size_t)k * nF * 3;
        int row = 0, found_contours = 0;
        for (int s0 = 0; s0 < S; ++s0) {                                      /* :130 */
            if (used[s0]) continue;
            /* Rows past nF would be out of bounds in the reference too (S_k + C_k > F needs every face to
             * cross); they are dropped here and in the CUDA path. */
            if (found_contours > 0) {                                         /* :135-140 separator */
                if (row < nF) out[3 * row] = out[3 * row + 1] = out[3 * row + 2] = -1e9f;
                ++row;
            }
            if (row < nF) memcpy(out + 3 * row, seg + 6 * s0, 3 * sizeof(float)); /* :143-146 seed p1 */
            ++row;
            used[s0] = 1;
            ++found_contours;
            float last[3] = {seg[6 * s0 + 3], seg[6 * s0 + 4], seg[6 * s0 + 5]}; /* :151-153 */
            int found = 1;
            while (found) {                                                   /* :157-203 */
                found = 0;
                for (int n = 0; n < S; ++n) {
                    if (used[n]) continue;
                    const float* p1 = seg + 6 * n;
                    const float* p2 = seg + 6 * n + 3;
                    float d1 = (last[0] - p1[0]) * (last[0] - p1[0]) + (last[1] - p1[1]) * (last[1] - p1[1]) +
                               (last[2] - p1[2]) * (last[2] - p1[2]);
                    float d2 = (last[0] - p2[0]) * (last[0] - p2[0]) + (last[1] - p2[1]) * (last[1] - p2[1]) +
                               (last[2] - p2[2]) * (last[2] - p2[2]);
                    if (d1 < thr) {
                        if (row < nF) memcpy(out + 3 * row, p2, 3 * sizeof(float));
                        memcpy(last, p2, 3 * sizeof(float));
                        ++row; used[n] = 1; found = 1; break;
                    } else if (d2 < thr) {
                        if (row < nF) memcpy(out + 3 * row, p1, 3 * sizeof(float));
                        memcpy(last, p1, 3 * sizeof(float));
                        ++row; used[n] = 1; found = 1; break;
                    }
                }
            }
        }
        if (found_contours > 0) {                                             /* :205-210 trailing separator */
            if (row < nF) out[3 * row] = out[3 * row + 1] = out[3 * row + 2] = -1e9f;
            ++row;
        }
        nums[k] = found_contours;                                             /* :212 */
    }
    free(seg); free(used);
    return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* get_fields_intersection_inter_slice (ext.cpp:20-46 -> intersection.cu:227-287, kernel :31-224)  */

static const float kOff[8][3] = {                                             /* intersection.cu:55-64 */
    {-0.5f, -0.5f, -0.5f}, {0.5f, -0.5f, -0.5f}, {-0.5f, 0.5f, -0.5f}, {0.5f, 0.5f, -0.5f},
    {-0.5f, -0.5f, 0.5f},  {0.5f, -0.5f, 0.5f},  {-0.5f, 0.5f, 0.5f},  {0.5f, 0.5f, 0.5f}};
static const int kEdge[12][2] = {                                             /* intersection.cu:67-71 */
    {0, 1}, {0, 2}, {1, 3}, {2, 3}, {4, 5}, {4, 6}, {5, 7}, {6, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

static int min_max8(const float* v, float* mn, float* mx) {                   /* intersection.cu:8-27 */
    *mn = v[0]; *mx = v[0];                                                   /* note: v[0] is not NaN-tested */
    for (int i = 1; i < 8; ++i) {
        float val = v[i];
        if (isnan(val)) return 0;
        if (val < *mn) *mn = val;
        if (val > *mx) *mx = val;
    }
    return 1;
}

/* mean edge-crossing point of one field in one voxel, intersection.cu:135-160 (and :163-199 twins).
 * nvcc (-fmad=true default) contracts `center + voxel_size*off` and `c1 + t*(c2-c1)` into FMAs. */
static int edge_mean(const float* f8, float iso, const float* ctr, float vs, float* acc) {
    int cnt = 0;
    acc[0] = acc[1] = acc[2] = 0.f;
    for (int e = 0; e < 12; ++e) {
        int c1 = kEdge[e][0], c2 = kEdge[e][1];
        float a = f8[c1], b = f8[c2];
        if (iso >= fminf(a, b) && iso <= fmaxf(a, b)) {                       /* :143-144 inclusive */
            float t = (iso - a) / (b - a);                                    /* :147 unguarded */
            for (int d = 0; d < 3; ++d) {
                float x1 = fmaf(vs, kOff[c1][d], ctr[d]);
                float x2 = fmaf(vs, kOff[c2][d], ctr[d]);
                acc[d] += fmaf(t, x2 - x1, x1);
            }
            ++cnt;
        }
    }
    return cnt;
}

static void voxel_cell_point(const float* s8, const float* a8, const float* b8, float s_iso, float a_iso,
                             float b_iso, const float* ctr, float vs, float* out) {
    float ps[3], pa[3], pb[3];
    int cs = edge_mean(s8, s_iso, ctr, vs, ps);
    int ca = edge_mean(a8, a_iso, ctr, vs, pa);
    int cb = edge_mean(b8, b_iso, ctr, vs, pb);
    for (int d = 0; d < 3; ++d) {                                             /* :203-213 */
        if (cs > 0) ps[d] /= (float)cs;
        if (ca > 0) pa[d] /= (float)ca;
        if (cb > 0) pb[d] /= (float)cb;
        out[d] = (float)((double)(ps[d] + pa[d] + pb[d]) / 3.0);              /* `/ 3.0` is a double division */
    }
}

/* Fields are [Dx,Dy,Dz,8] corner-expanded, centers [Dx,Dy,Dz,3].  Outputs: mask [Ns,N1,N2] (0/1 bytes),
 * idx [Ns,N1,N2,3] int16 (-1 fill, intersection.cu:247), pts [Ns,N1,N2,3] (zero fill, :248). */
int oracle_fields_intersection(const float* s, const float* l1, const float* l2, const float* s_iso,
                               const float* l1_iso, const float* l2_iso, const float* centers, int Dx, int Dy,
                               int Dz, int Ns, int N1, int N2, uint8_t* mask, int16_t* idx, float* pts,
                               int winner_rule) {
    size_t cells = (size_t)Ns * N1 * N2;
    memset(mask, 0, cells);
    for (size_t i = 0; i < cells * 3; ++i) idx[i] = -1;
    memset(pts, 0, cells * 3 * sizeof(float));
    const float vs = (float)(2.0 / (double)(float)Dx);                        /* :54 */
    for (int x = 0; x < Dx; ++x)
        for (int y = 0; y < Dy; ++y)
            for (int z = 0; z < Dz; ++z) {
                size_t v = ((size_t)x * Dy + y) * Dz + z;
                float amn, amx, bmn, bmx, smn, smx;
                if (!min_max8(l1 + 8 * v, &amn, &amx)) continue;              /* :79-91 order l1, l2, slice */
                if (!min_max8(l2 + 8 * v, &bmn, &bmx)) continue;
                if (!min_max8(s + 8 * v, &smn, &smx)) continue;
                for (int i1 = 0; i1 < N1; ++i1) {                             /* :95-116 */
                    float ai = l1_iso[i1];
                    if (ai < amn || ai > amx) continue;
                    for (int i2 = 0; i2 < N2; ++i2) {
                        float bi = l2_iso[i2];
                        if (bi < bmn || bi > bmx) continue;
                        for (int is = 0; is < Ns; ++is) {
                            float si = s_iso[is];
                            if (si < smn || si > smx) continue;
                            size_t c = ((size_t)is * N1 + i1) * N2 + i2;
                            if (winner_rule == 0 && mask[c]) continue;
                            idx[3 * c] = (int16_t)is; idx[3 * c + 1] = (int16_t)i1; idx[3 * c + 2] = (int16_t)i2;
                            mask[c] = 1;
                            voxel_cell_point(s + 8 * v, l1 + 8 * v, l2 + 8 * v, si, ai, bi, centers + 3 * v, vs,
                                             pts + 3 * c);
                        }
                    }
                }
            }
    return 0;
}

/* Race-tolerant check of a (mask, pts) pair produced by ANY schedule of the reference kernel.  The reference
 * writes the three coordinates of a cell with three plain stores (intersection.cu:216-218), so when several
 * voxels hit one cell the result may even be TORN (x from one voxel, y/z from another): a cell passes if every
 * coordinate equals, within `tol`, that coordinate of some voxel that hits the cell (NaN matches NaN).
 * Returns the number of failing cells: cells with an unexplained coordinate, plus cells whose mask disagrees
 * with "some voxel hits". */
long oracle_intersection_unmatched(const float* s, const float* l1, const float* l2, const float* s_iso,
                                   const float* l1_iso, const float* l2_iso, const float* centers, int Dx, int Dy,
                                   int Dz, int Ns, int N1, int N2, const uint8_t* mask, const float* pts,
                                   float tol) {
    size_t cells = (size_t)Ns * N1 * N2;
    uint8_t* hit = (uint8_t*)calloc(cells, 1);
    uint8_t* ok = (uint8_t*)calloc(cells, 1);   /* bit d set: coordinate d explained */
    if (!hit || !ok) { free(hit); free(ok); return -1; }
    const float vs = (float)(2.0 / (double)(float)Dx);
    for (int x = 0; x < Dx; ++x)
        for (int y = 0; y < Dy; ++y)
            for (int z = 0; z < Dz; ++z) {
                size_t v = ((size_t)x * Dy + y) * Dz + z;
                float amn, amx, bmn, bmx, smn, smx;
                if (!min_max8(l1 + 8 * v, &amn, &amx)) continue;
                if (!min_max8(l2 + 8 * v, &bmn, &bmx)) continue;
                if (!min_max8(s + 8 * v, &smn, &smx)) continue;
                for (int i1 = 0; i1 < N1; ++i1) {
                    float ai = l1_iso[i1];
                    if (ai < amn || ai > amx) continue;
                    for (int i2 = 0; i2 < N2; ++i2) {
                        float bi = l2_iso[i2];
                        if (bi < bmn || bi > bmx) continue;
                        for (int is = 0; is < Ns; ++is) {
                            float si = s_iso[is];
                            if (si < smn || si > smx) continue;
                            size_t c = ((size_t)is * N1 + i1) * N2 + i2;
                            hit[c] = 1;
                            if (ok[c] == 7) continue;
                            float p[3];
                            voxel_cell_point(s + 8 * v, l1 + 8 * v, l2 + 8 * v, si, ai, bi, centers + 3 * v, vs, p);
                            for (int d = 0; d < 3; ++d) {
                                float q = pts[3 * c + d];
                                if ((isnan(p[d]) && isnan(q)) || fabsf(p[d] - q) <= tol) ok[c] |= (uint8_t)(1 << d);
                            }
                        }
                    }
                }
            }
    long bad = 0;
    for (size_t c = 0; c < cells; ++c) {
        if ((mask[c] != 0) != (hit[c] != 0)) ++bad;
        else if (hit[c] && ok[c] != 7) ++bad;
    }
    free(hit); free(ok);
    return bad;
}
