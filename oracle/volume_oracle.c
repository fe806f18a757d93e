/* ORACLE -- test infrastructure, NOT product code.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load this.
 *
 * Single-thread C restatement of the volume-side producers:
 *
 *   oracle_grid_queries      sdf_meshing.py:46-65 gen_voxle_queries, including the true-division shear of :57-58
 *                            (pinned: tests/golden/siren_golden.npz holds the reference's own output)
 *   oracle_corner_fields     toolpath_utils.py:986-1019 extract_corner_fields
 *   oracle_voxel_centers     toolpath_utils.py:1021-1033 get_voxel_centers
 *                            (both pinned: tests/golden/volume_golden.npz, recorded from the reference's functions)
 *   oracle_marching_cubes    iso-surface of a sample grid in the conventions of scripts/gen_mc_table.py.
 *                            PARITY UNPINNED: the reference calls skimage.measure.marching_cubes (scikit-image ^0.25.2,
 *                            pyproject.toml:16; call sites sdf_meshing.py:399, utils.py:322), which is neither in
 *                            /root/reference nor installed here, and the reference has no test or golden mesh.  What IS
 *                            pinned is what every marching-cubes variant shares: one welded vertex per sign-changing
 *                            grid edge at the linear interpolation of the level (the set skimage's Lewiner variant
 *                            produces too, except for its extra in-cell vertices in ambiguous cubes), plus the
 *                            properties the tests assert (closed oriented manifold, Euler characteristic, |f(v)-level|).
 *
 * Serial sweep in linear point order (z fastest): vertices are numbered by owner point then axis, triangles by cube
 * then table order -- the order the CUDA path reproduces, so comparisons are bit for bit.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (no FMA contraction; every operation rounds to float32).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../inf-3dp_b200/csrc/mc_table.inc"

void oracle_grid_queries(int N, float cube_size, long long start, long long count, float* out) {
    const float voxel_size = (float)((double)cube_size * 2.0 / (double)(N - 1));
    const float origin = -cube_size, fN = (float)N;
    for (long long i = 0; i < count; ++i) {
        const long long idx = start + i;
        const float iz = (float)(idx % N);
        const float q1 = (float)idx / fN;                 /* LongTensor / int -> float32 true division */
        const float iy = fmodf(q1, fN);
        const float q2 = q1 / fN;
        const float ix = fmodf(q2, fN);
        float x = ix * voxel_size; x = x + origin;
        float y = iy * voxel_size; y = y + origin;
        float z = iz * voxel_size; z = z + origin;
        out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
    }
}

void oracle_corner_fields(const float* grid, int Gx, int Gy, int Gz, float* out /* [Gx-1,Gy-1,Gz-1,8] */) {
    long long o = 0;
    for (int x = 0; x + 1 < Gx; ++x)
        for (int y = 0; y + 1 < Gy; ++y)
            for (int z = 0; z + 1 < Gz; ++z)
                for (int c = 0; c < 8; ++c)
                    out[o++] = grid[((long long)(x + (c & 1)) * Gy + (y + ((c >> 1) & 1))) * Gz + (z + ((c >> 2) & 1))];
}

void oracle_voxel_centers(int N, float* out /* [N-1,N-1,N-1,3] */) {
    const float vs = (float)(2.0 / (double)(N - 1));
    long long o = 0;
    for (int x = 0; x + 1 < N; ++x)
        for (int y = 0; y + 1 < N; ++y)
            for (int z = 0; z + 1 < N; ++z) {
                const int id[3] = {x, y, z};
                for (int d = 0; d < 3; ++d) {
                    float c = (float)id[d] + 0.5f;
                    c = c * vs;
                    c = c + (-1.0f);
                    out[o++] = c;
                }
            }
}

/* ------------------------------------------------------------------------------------------------------------------ */
typedef struct { const float* v; int G[3]; long long s[3]; float level; } Grid;

static int fin(float x) { return fabsf(x) <= 3.402823466e38f; }
static float at(const Grid* g, int x, int y, int z) { return g->v[x * g->s[0] + y * g->s[1] + z * g->s[2]]; }

static int cube_valid(const Grid* g, int x, int y, int z) {
    if (x < 0 || y < 0 || z < 0 || x + 1 >= g->G[0] || y + 1 >= g->G[1] || z + 1 >= g->G[2]) return 0;
    for (int c = 0; c < 8; ++c)
        if (!fin(at(g, x + (c & 1), y + ((c >> 1) & 1), z + ((c >> 2) & 1)))) return 0;
    return 1;
}

static int cube_case(const Grid* g, int x, int y, int z) {
    if (!cube_valid(g, x, y, z)) return 0;
    int cs = 0;
    for (int c = 0; c < 8; ++c)
        if (at(g, x + (c & 1), y + ((c >> 1) & 1), z + ((c >> 2) & 1)) < g->level) cs |= 1 << c;
    return cs;
}

/* vertex on the edge from (x,y,z) along axis a? */
static int edge_vertex(const Grid* g, int x, int y, int z, int a) {
    int q[3] = {x, y, z};
    if (q[a] + 1 >= g->G[a]) return 0;
    q[a] += 1;
    const float va = at(g, x, y, z), vb = at(g, q[0], q[1], q[2]);
    if (!fin(va) || !fin(vb) || ((va < g->level) == (vb < g->level))) return 0;
    const int u1 = (a == 0) ? 1 : 0, u2 = (a == 2) ? 1 : 2;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            int c[3] = {x, y, z};
            c[u1] -= i; c[u2] -= j;
            if (cube_valid(g, c[0], c[1], c[2])) return 1;
        }
    return 0;
}

static float grad_axis(const Grid* g, const int* p, int a, float h) {
    const float c = at(g, p[0], p[1], p[2]);
    int lo_i[3] = {p[0], p[1], p[2]}, hi_i[3] = {p[0], p[1], p[2]};
    lo_i[a] -= 1; hi_i[a] += 1;
    const int has_lo = p[a] > 0, has_hi = p[a] + 1 < g->G[a];
    const float lo = has_lo ? at(g, lo_i[0], lo_i[1], lo_i[2]) : 0.f, hi = has_hi ? at(g, hi_i[0], hi_i[1], hi_i[2]) : 0.f;
    const int ok_lo = has_lo && fin(lo), ok_hi = has_hi && fin(hi);
    if (ok_lo && ok_hi) { float d = hi - lo; d = d * 0.5f; return d / h; }
    if (ok_hi) { float d = hi - c; return d / h; }
    if (ok_lo) { float d = c - lo; return d / h; }
    return 0.f;
}

/* Two calls: with verts == NULL only the counts are returned. */
int oracle_marching_cubes(const float* volume, int Gx, int Gy, int Gz, float level, const float* spacing, int descent,
                          float* verts, int32_t* faces, float* normals, float* values, long long* counts) {
    Grid g = {volume, {Gx, Gy, Gz}, {(long long)Gy * Gz, Gz, 1}, level};
    const long long points = (long long)Gx * Gy * Gz;
    int32_t* first = NULL;      /* first vertex id of every point, flags in a second array */
    unsigned char* flags = NULL;
    if (verts) {
        first = (int32_t*)malloc((size_t)points * sizeof(int32_t));
        flags = (unsigned char*)calloc((size_t)points, 1);
        if (!first || !flags) { free(first); free(flags); return 1; }
    }
    long long nv = 0, nf = 0;
    for (int x = 0; x < Gx; ++x)
        for (int y = 0; y < Gy; ++y)
            for (int z = 0; z < Gz; ++z) {
                const long long p = ((long long)x * Gy + y) * Gz + z;
                if (verts) first[p] = (int32_t)nv;
                for (int a = 0; a < 3; ++a) {
                    if (!edge_vertex(&g, x, y, z, a)) continue;
                    if (verts) {
                        flags[p] |= (unsigned char)(1 << a);
                        int q[3] = {x, y, z};
                        q[a] += 1;
                        const int p0[3] = {x, y, z};
                        const float va = at(&g, x, y, z), vb = at(&g, q[0], q[1], q[2]);
                        float num = level - va, den = vb - va;
                        const float t = num / den;
                        for (int d = 0; d < 3; ++d) {
                            float c = (float)p0[d];
                            if (d == a) c = c + t;
                            verts[3 * nv + d] = c * spacing[d];
                        }
                        if (values) values[nv] = va > vb ? va : vb;
                        if (normals) {
                            float n[3];
                            for (int d = 0; d < 3; ++d) {
                                const float g0 = grad_axis(&g, p0, d, spacing[d]), g1 = grad_axis(&g, q, d, spacing[d]);
                                float diff = g1 - g0;
                                diff = t * diff;
                                n[d] = g0 + diff;
                            }
                            float s2 = n[0] * n[0];
                            float t1 = n[1] * n[1];
                            s2 = s2 + t1;
                            t1 = n[2] * n[2];
                            s2 = s2 + t1;
                            const float len = sqrtf(s2);
                            for (int d = 0; d < 3; ++d) {
                                float r = 0.f;
                                if (len > 0.f && fin(len)) { r = n[d] / len; if (descent) r = -r; }
                                normals[3 * nv + d] = r;
                            }
                        }
                    }
                    ++nv;
                }
            }
    for (int x = 0; x < Gx; ++x)
        for (int y = 0; y < Gy; ++y)
            for (int z = 0; z < Gz; ++z) {
                const int cs = cube_case(&g, x, y, z);
                const int nt = INF3DP_MC_NTRI[cs];
                for (int k = 0; k < nt; ++k) {
                    if (faces) {
                        int id[3];
                        for (int j = 0; j < 3; ++j) {
                            const int e = INF3DP_MC_TRI[cs][3 * k + j];
                            const int a = e >> 2, kk = e & 3;
                            const int u1 = (a == 0) ? 1 : 0, u2 = (a == 2) ? 1 : 2;
                            int q[3] = {x, y, z};
                            q[u1] += kk & 1; q[u2] += kk >> 1;
                            const long long p = ((long long)q[0] * Gy + q[1]) * Gz + q[2];
                            int rank = 0;
                            for (int b = 0; b < a; ++b) rank += (flags[p] >> b) & 1;
                            if (!((flags[p] >> a) & 1)) { free(first); free(flags); return 2; }   /* table/ownership bug */
                            id[j] = first[p] + rank;
                        }
                        faces[3 * nf] = id[0];
                        faces[3 * nf + 1] = descent ? id[2] : id[1];
                        faces[3 * nf + 2] = descent ? id[1] : id[2];
                    }
                    ++nf;
                }
            }
    counts[0] = nv; counts[1] = nf;
    free(first); free(flags);
    return 0;
}
