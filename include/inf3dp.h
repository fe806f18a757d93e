/* inf3dp.h -- C ABI of the B200-native INF-3DP hot path (libinf3dp.so, sm_100a only).
 *
 * Plain pointers and sizes, no torch types.  Every pointer argument is a DEVICE pointer unless its name ends
 * in `_host`.  Inputs are borrowed and never written; outputs and workspaces are allocated by the caller
 * (the Python shim allocates them as torch tensors); kernels never allocate.  All entry points are
 * asynchronous on `stream` unless stated, re-entrant, and return 0 on success or an INF3DP_E* code, with a
 * human-readable message available from inf3dp_last_error() (thread-local).  There is no CPU fallback.
 *
 * Each entry point names the reference interface it replaces (paths relative to the INF-3DP repository).
 */
#ifndef INF3DP_H_
#define INF3DP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* inf3dp_stream_t; /* == cudaStream_t */

enum {
    INF3DP_OK = 0,
    INF3DP_EINVAL = 1,      /* bad argument (shape, null pointer, unsupported configuration) */
    INF3DP_ECUDA = 2,       /* a CUDA runtime call or kernel launch failed                   */
    INF3DP_EWORKSPACE = 3,  /* workspace too small (see the *_workspace_bytes query)          */
    INF3DP_EOVERFLOW = 4    /* data-dependent capacity exceeded (reported by *_status)       */
};

enum { INF3DP_ACT_SINE = 0, INF3DP_ACT_RELU = 1 };

/* Execution engine of inf3dp_siren_jet. */
enum {
    INF3DP_MODE_AUTO = 0,       /* TC_PRECISE when the configuration supports it, else SIMT           */
    INF3DP_MODE_SIMT = 1,       /* FP32 CUDA-core kernel: any supported configuration, orders 0..2     */
    INF3DP_MODE_TC_PRECISE = 2, /* tcgen05 tensor cores, FP16 hi/lo split operands (3 MMAs), FP32 acc. */
    INF3DP_MODE_TC_FAST = 3,    /* tcgen05 tensor cores, single FP16 operands (does NOT meet 0.1 deg)  */
    INF3DP_MODE_TC_REVERSE = 4  /* tcgen05 CTA pairs, single-output sine networks, FP16 hi/lo split operands: the
                                   forward sweep alone (order 0), reverse-mode value+gradient (order 1) or
                                   forward-over-reverse value+gradient+Hessian (order 2); orders 1 and 2 need the
                                   workspace of inf3dp_siren_jet_ws                                           */
};

const char* inf3dp_last_error(void);
int inf3dp_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * SIREN field network.  Replaces modules.py:143-160 (SingleBVPNet.forward -> FCBlock -> BatchLinear + Sine)
 * together with what diff_operators.py:128-132 (gradient) and :27-44 (hessian3d) extract from it by autograd.
 *
 * Network: in_features -> hidden (act) -> [hidden -> hidden (act)] x num_hidden_layers -> out_features (linear),
 * act = sin(omega0 * z) (modules.py:34) or max(z, 0) (exp_scripts/train_levels.py:35-51).
 * Supported: hidden == 256, in_features in {3, 4}, 1 <= out_features <= 64, num_hidden_layers >= 0.
 * ------------------------------------------------------------------------------------------------------- */

/* Forget the host-side header mirror of a packed blob.  Call before the blob's memory is freed or overwritten by anything
 * other than inf3dp_siren_pack_weights: the library caches blob headers by device address (so that queries never read the
 * device synchronously), and a different blob later placed at the same address must not meet a stale entry.  Never fails. */
int inf3dp_siren_release(const void* packed_weights);

/* Size of the packed-weights blob for a configuration (0 if unsupported). */
size_t inf3dp_siren_packed_bytes(int in_features, int hidden, int num_hidden_layers, int out_features);

/* Pack a state_dict once: folds omega0 into the sine layers in FP32 and writes the operand images every
 * engine needs (FP32 transposed weights for SIMT; FP16 hi/lo UMMA shared-memory images for tcgen05).
 * weights_host / biases_host: host arrays of (num_hidden_layers + 2) DEVICE pointers, layer i of shape
 * [out_i, in_i] row-major / [out_i] -- the tensors behind the reference keys net.net.{i}.0.weight / .bias. */
int inf3dp_siren_pack_weights(const float* const* weights_host, const float* const* biases_host, int in_features,
                              int hidden, int num_hidden_layers, int out_features, int activation, float omega0,
                              void* packed, size_t packed_bytes, inf3dp_stream_t stream);

/* Batched field query with forward-mode derivatives ("jet").
 *   x [n, in_features] -> y [n, out];  order >= 1: J [n, out, 3] = dy/dx;  order == 2: H [n, out, 3, 3].
 * Derivatives are taken with respect to the first three input features.  J / H may be NULL when order is lower.
 * mode: INF3DP_MODE_*.  TC modes need the flagship configuration (sine, in 3, hidden 256, 3 hidden layers,
 * out <= 4) with order <= 1, or out == 1 with order == 2 (second-order jet on tensor-core CTA pairs: 10 jet rows per
 * point), or the ReLU classifier in 4 / one hidden layer / out <= 16 with order 0 (x must then be 16-byte aligned), and
 * return INF3DP_EINVAL otherwise; AUTO falls back to SIMT for everything else. */
int inf3dp_siren_jet(const void* packed, const float* x, int64_t n, int order, float* y, float* J, float* H,
                     int mode, inf3dp_stream_t stream);

/* Same query with a caller-provided scratch buffer of inf3dp_siren_jet_workspace_bytes() bytes (device memory,
 * contents irrelevant, must not be shared by launches that may run concurrently).  With a workspace, AUTO runs
 * value+gradient queries of single-output sine networks on the reverse-mode engine (INF3DP_MODE_TC_REVERSE): the
 * same forward + backward sweep the reference's autograd performs (diff_operators.py:128-132), fused in one kernel,
 * half the tensor work of the forward-mode jet.  The scratch holds the cosines of two hidden layers per resident
 * tile (256 KB per SM) between the sweeps.  Order-2 queries of such networks (diff_operators.py:6-44 hessian / hessian3d)
 * then run forward-over-reverse on the same engine -- the point and its three tangents d/dx_i through the same forward +
 * backward sweep, 32 points per tile, Hessian rows symmetrised -- instead of the 10-row second-order forward jet that
 * INF3DP_MODE_TC_PRECISE (and AUTO without a workspace) selects.  Order-0 queries of such networks run the forward sweep of
 * the same engine (no workspace needed); the value is bit-identical to the one an order-1 query returns. */
size_t inf3dp_siren_jet_workspace_bytes(void);
/* Process-wide cap on the SMs the persistent SIREN engines occupy (0 = all, the default).  A multi-GPU caller that overlaps
 * the all-gather of chunk i with the query of chunk i+1 leaves a few SMs to the NCCL kernels (inf3dp/dist.py). */
int inf3dp_siren_set_sm_limit(int max_sms);
int inf3dp_siren_jet_ws(const void* packed, const float* x, int64_t n, int order, float* y, float* J, float* H,
                        int mode, void* workspace, size_t workspace_bytes, inf3dp_stream_t stream);

/* Unit normals: normals = grad / max(|grad|, eps), rows of 3 (torch.nn.functional.normalize as applied after
 * field_query_with_grads, toolpath_utils.py:180-181).  In-place (normals == grad) is allowed. */
int inf3dp_unit_normals(const float* grad, int64_t n, float eps, float* normals, inf3dp_stream_t stream);

/* Value + gradient of a single-output field network (the reverse-mode engine of inf3dp_siren_jet_ws, order 1) written as
 * packed rows [n,4] = (value, d/dx, d/dy, d/dz), 16-byte aligned: the layout of the reference's host result
 * (utils.py:249-253 concatenates fields and grads) and of the multi-GPU all-gather, one 16-byte store per point.
 * inf3dp_unit_normals_rows is inf3dp_unit_normals on such rows. */
int inf3dp_siren_value_grad_rows(const void* packed, const float* x, int64_t n, float* rows, void* workspace,
                                 size_t workspace_bytes, inf3dp_stream_t stream);
int inf3dp_unit_normals_rows(const float* rows, int64_t n, float eps, float* normals, inf3dp_stream_t stream);
/* The same query FUSED WITH THE MULTI-GPU ALL-GATHER of its rows (points sharded, weights replicated,
 * one final all-gather -- the reference has no distributed code).  The engine's epilogue stores every row directly into
 * the gathered buffers of the ranks of the job, over NVLink, tile by tile while the tensor cores work on the next tile:
 *   multicast_rows != NULL  one multimem.st per point to an NVLink multicast address (the NVSwitch replicates it to every
 *                           rank, this one included); peer_rows is ignored
 *   otherwise               one 16-byte store per point and peer to each of peer_rows[0..n_peers) (1 <= n_peers <= 8;
 *                           HOST array of peer-mapped DEVICE pointers, e.g. CUDA IPC / symmetric-memory mappings)
 * Each pointer addresses the destination of row 0 of THIS launch (the caller applies rank and chunk offsets).  The
 * rows become visible to the other ranks after this launch has completed AND the ranks have synchronised (any
 * cross-rank barrier with system-scope release/acquire on the same stream); no SM is set aside for a communication kernel. */
int inf3dp_siren_value_grad_rows_peers(const void* packed, const float* x, int64_t n, float* const* peer_rows, int n_peers,
                                       float* multicast_rows, void* workspace, size_t workspace_bytes,
                                       inf3dp_stream_t stream);

/* Curvature tail.  Replaces the per-point Python loop of diff_operators.py:167-220 (min_max_curvs_and_vectors)
 * and construct_tangent_basis (:280-298): grad [n,3], hess [n,3,3] -> kappa [n,2] (min,max), dirs [n,2,3]. */
int inf3dp_principal_curvatures(const float* grad, const float* hess, int64_t n, float* kappa, float* dirs,
                                inf3dp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Isocontour extraction.  Replaces contour_cuda/ext.cpp:5-17 extract_isocontours
 * (-> isocontours.cu:216-293: extract_contours_kernel :16-93 + merge_segments_kernel :99-213).
 *   vertices [nV,3] f32, faces [nF,3] i32, fields [nV] f32, iso_values [K] f32
 *   -> contours_merged [K, nF, 3] f32 (dense, zero padded, loops separated by (-1e9,-1e9,-1e9) rows),
 *      contour_nums [K] i32.  Both outputs are fully written by the call (no pre-zeroing needed).
 * Deterministic: equals the reference run with segments slotted in ascending face order.
 * ------------------------------------------------------------------------------------------------------- */
size_t inf3dp_extract_isocontours_workspace_bytes(int nV, int nF, int K);
/* Same, but with a segment pool of at least `min_segments` (use after INF3DP_EOVERFLOW; a larger workspace
 * passed to inf3dp_extract_isocontours is used automatically). */
size_t inf3dp_extract_isocontours_workspace_bytes_segments(int nV, int nF, int K, int64_t min_segments);
int inf3dp_extract_isocontours(const float* vertices, const int32_t* faces, const float* fields,
                               const float* iso_values, int nV, int nF, int K, float* contours_merged,
                               int32_t* contour_nums, void* workspace, size_t workspace_bytes,
                               inf3dp_stream_t stream);
/* Synchronises `stream` and reports data-dependent failures of the last call that used `workspace`
 * (INF3DP_EOVERFLOW when the segment pool was too small).  seg_counts_host: optional host [K] array receiving
 * the per-isovalue segment counts S_k (the reference's internal pairpoint_nums). */
int inf3dp_extract_isocontours_status(const void* workspace, int K, int32_t* seg_counts_host,
                                      inf3dp_stream_t stream);
/* Synchronises `stream` and returns the total number of segments the last call that used `workspace` produced (or
 * would have produced: after INF3DP_EOVERFLOW this is the pool size to pass to
 * inf3dp_extract_isocontours_workspace_bytes_segments for the retry); -1 on a CUDA error. */
int64_t inf3dp_extract_isocontours_required_segments(const void* workspace, int K, inf3dp_stream_t stream);
/* Threading: the dense entry point forks its zero fill and its chain onto ONE internal, process-wide, highest-priority
 * helper stream per device (created on first use) and joins it back into `stream` before returning, on every exit path;
 * the fork / join events come from an internal pool.  Concurrent calls on different streams are safe (their chains
 * serialise on the helper stream); the compact entry point uses `stream` only. */

/* Compact variant: the same row streams without the dense zero-padded [K, nF, 3] contract.
 * Isovalue k's rows (loops separated and terminated by (-1e9,-1e9,-1e9) rows, exactly the first row_counts[k] rows of
 * contours_merged[k]) are written at rows[row_offsets[k] ...]; row_counts[k] = S_k + C_k, contour_nums[k] = C_k.
 * `rows` must hold inf3dp_extract_isocontours_compact_rows(...) rows of 3 floats (twice the segment pool of the
 * workspace); nothing else of it is written.  Replaces the dense device-to-host copy + host scan of
 * toolpath_utils.py:38-134 (split_merged_contours) by a transfer of the used rows only. */
int64_t inf3dp_extract_isocontours_compact_rows(int nV, int nF, int K, size_t workspace_bytes);
int inf3dp_extract_isocontours_compact(const float* vertices, const int32_t* faces, const float* fields,
                                       const float* iso_values, int nV, int nF, int K, float* rows,
                                       int64_t rows_capacity, int64_t* row_offsets, int32_t* row_counts,
                                       int32_t* contour_nums, void* workspace, size_t workspace_bytes,
                                       inf3dp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Three-field voxel intersection.  Replaces contour_cuda/ext.cpp:20-46 get_fields_intersection_inter_slice
 * (-> intersection.cu:227-287, kernel :31-224).
 *   slice/lattice1/lattice2 fields [Dx,Dy,Dz,8] f32 (corner expanded), iso arrays [Ns],[N1],[N2] f32,
 *   voxel_centers [Dx,Dy,Dz,3] f32
 *   -> inter_mask [Ns,N1,N2] u8 (0/1), isovalue_indices [Ns,N1,N2,3] i16 (-1 where unset),
 *      intersections [Ns,N1,N2,3] f32.  All three fully written by the call.
 * Deterministic: where several voxels hit one cell the LOWEST linear voxel index wins
 * (the reference is last-writer-wins, i.e. any of them).
 * ------------------------------------------------------------------------------------------------------- */
size_t inf3dp_fields_intersection_workspace_bytes(int Dx, int Dy, int Dz, int Ns, int N1, int N2);
int inf3dp_fields_intersection(const float* slice_fields, const float* lattice1_fields, const float* lattice2_fields,
                               const float* slice_iso, const float* lattice1_iso, const float* lattice2_iso,
                               const float* voxel_centers, int Dx, int Dy, int Dz, int Ns, int N1, int N2,
                               uint8_t* inter_mask, int16_t* isovalue_indices, float* intersections,
                               void* workspace, size_t workspace_bytes, inf3dp_stream_t stream);


/* The same operation on the SAMPLE GRIDS themselves.  Replaces, in one call, the reference's host-side preparation
 * toolpath_utils.py:986-1019 extract_corner_fields (8x corner expansion of each [N,N,N] grid) and :1021-1033
 * get_voxel_centers, followed by ext.cpp:20-46: grids [Gx,Gy,Gz] f32 (voxels = (Gx-1)(Gy-1)(Gz-1), corner c of voxel
 * (x,y,z) = grid[x+(c&1), y+((c>>1)&1), z+((c>>2)&1)], centre = (index + 0.5) * float(2/(Gx-1)) - 1 rounded as
 * torch does), same outputs, bit-identical to inf3dp_fields_intersection on the expanded tensors.
 * Workspace: inf3dp_fields_intersection_workspace_bytes(Gx-1, Gy-1, Gz-1, Ns, N1, N2). */
int inf3dp_fields_intersection_grid(const float* slice_grid, const float* lattice1_grid, const float* lattice2_grid,
                                    const float* slice_iso, const float* lattice1_iso, const float* lattice2_iso,
                                    int Gx, int Gy, int Gz, int Ns, int N1, int N2, uint8_t* inter_mask,
                                    int16_t* isovalue_indices, float* intersections, void* workspace,
                                    size_t workspace_bytes, inf3dp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Query grid.  Replaces sdf_meshing.py:46-65 gen_voxle_queries for an index range: out [count,3] f32 holds rows
 * [start, start+count) of the reference's [N^3,3] tensor, bit-identical INCLUDING the true-division shear of
 * :57-58.  A grid sweep therefore never materialises the 12*N^3-byte coordinate tensor.
 * ------------------------------------------------------------------------------------------------------- */
int inf3dp_grid_queries(int N, float cube_size, int64_t start, int64_t count, float* out, inf3dp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Marching cubes on a sample grid.  Replaces the host call skimage.measure.marching_cubes(volume, level, spacing)
 * at sdf_meshing.py:399 (convert_sdf_samples_to_ply) and utils.py:322 (gen_iso_layers); same return values:
 *   verts [nv,3] f32 (index coordinates * spacing), faces [nf,3] i32, normals [nv,3] f32 (normalised grid gradient),
 *   values [nv] f32 (larger of the two samples of the vertex's edge).
 * One welded vertex per sign-changing grid edge; cubes with a non-finite corner are skipped (NaN-masked volumes,
 * utils.py:308); `descent` != 0 orients triangles and normals towards decreasing values (skimage's default
 * gradient_direction), 0 towards increasing values.  Deterministic output order (linear sample order, z fastest).
 * Two phases because the sizes are data dependent:
 *   _count  classifies, leaves per-tile counts/offsets in the workspace, SYNCHRONISES the stream and returns
 *           counts_host[0] = nv, counts_host[1] = nf
 *   _emit   writes the arrays (normals / values may be NULL); needs the workspace left by _count for the same
 *           volume and level.
 * ------------------------------------------------------------------------------------------------------- */
size_t inf3dp_marching_cubes_workspace_bytes(int Gx, int Gy, int Gz);
int inf3dp_marching_cubes_count(const float* volume, int Gx, int Gy, int Gz, float level, void* workspace,
                                size_t workspace_bytes, int64_t* counts_host, inf3dp_stream_t stream);
int inf3dp_marching_cubes_emit(const float* volume, int Gx, int Gy, int Gz, float level, float spacing_x,
                               float spacing_y, float spacing_z, int descent, float* verts, int32_t* faces,
                               float* normals, float* values, int64_t n_verts, int64_t n_faces, void* workspace,
                               size_t workspace_bytes, inf3dp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * TV-SDF collision query: the fused predicate / compaction stages between the field sweeps.
 * Replaces the mask / nonzero / gather / where glue of level_collision.CollTVSDF.forward (level_collision.py:73-237) and
 * the posing of utils.transform_nozzle_pcd_with_quaternion (utils.py:612-621); the sweeps themselves are
 * inf3dp_siren_jet_ws / inf3dp_siren_value_grad_rows launches on the compact lists these stages write.
 * Point ids: p = waypoint * M + nozzle point (< 2^30 per call).  counters: int32[8] on the device
 * ([0] cube list, [1] inside list, [2] recheck list, [3] points whose result carries the SDF gradient); inf3dp_coll_pose
 * zeroes them, the caller reads a count back before it launches the next sweep.
 * Outputs ("out surface", either or both): per point depth_p [P] / mask_p [P] / grads_p [P,3] (the CollTVSDF surface; the
 * caller pre-zeroes them) and / or per waypoint depth_w [Nw] (max depth as IEEE bits, pre-zeroed) / hit_w [Nw].
 * ------------------------------------------------------------------------------------------------------- */
int inf3dp_coll_pose(const float* waypts, const float* quats /* [Nw,4] or NULL */, const float* nozzle_or_posed_cloud,
                     int64_t Nw, int M, float base_th, float* cube_x, int32_t* cube_id, int32_t* counters, float* depth_p,
                     uint8_t* mask_p, float* grads_p, uint32_t* depth_w, uint8_t* hit_w, inf3dp_stream_t stream);
int inf3dp_coll_inside(const float* sdf, int64_t n, const float* cube_x, const int32_t* cube_id, float* in_x, int32_t* in_id,
                       float* in_sdf, int32_t* counters, int M, float* depth_p, uint8_t* mask_p, float* grads_p,
                       uint32_t* depth_w, uint8_t* hit_w, inf3dp_stream_t stream);
int inf3dp_coll_pack4(const float* x3, const float* values, int value_stride, int64_t n, float* x4, inf3dp_stream_t stream);
int inf3dp_coll_predicate(const float* slice_rows, const float* logits, int C, int64_t n, const float* in_x,
                          const int32_t* in_id, const float* in_sdf, const int32_t* levels_w, const float* maxslices, int M,
                          float* re_x, int32_t* re_src, float* re_dist, int32_t* grad_src, int want_grads, int32_t* counters,
                          float* depth_p, uint8_t* mask_p, float* grads_p, uint32_t* depth_w, uint8_t* hit_w,
                          inf3dp_stream_t stream);
int inf3dp_coll_push(const float* slice_rows_at_push1, int64_t n, const float* re_x, const float* re_dist, float* push_x,
                     inf3dp_stream_t stream);
int inf3dp_coll_alter(const float* slice_at_push, const float* logits_at_push, int C, int64_t n, const int32_t* re_src,
                      const float* re_dist, const float* slice_rows, const int32_t* in_id, const float* in_sdf,
                      const int32_t* levels_w, const float* maxslices, int M, int32_t* grad_src, int want_grads,
                      int32_t* counters, float* depth_p, uint8_t* mask_p, float* grads_p, uint32_t* depth_w, uint8_t* hit_w,
                      inf3dp_stream_t stream);
int inf3dp_coll_gather_x(const int32_t* grad_src, int64_t n, const float* in_x, float* gx, inf3dp_stream_t stream);
int inf3dp_coll_scatter_grads(const float* sdf_rows, int64_t n, const int32_t* grad_src, const int32_t* in_id, float* grads_p,
                              inf3dp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* INF3DP_H_ */
