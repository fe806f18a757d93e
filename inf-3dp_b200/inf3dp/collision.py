"""Device-resident TV-SDF collision evaluation of nozzle point clouds against the printed part.

Reference: level_collision.CollTVSDF (level_collision.py:68-255, paper section 5.1) and its wrapper
collision_evaluation (:259-281).  The reference runs the same sequence through utils.field_query_with_grads /
level_querying_with_confidence, which build an autograd graph per chunk and bounce every intermediate through host
tensors (utils.py:232-303); here the five field sweeps are launches of the fused kernels on device-resident,
mask-compacted point sets.  Same predicates, same Eq. 27, same two-step push and the same
backward (negated, normalised saved directions times the upstream gradient).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .query import _net_of


def _value_grad(decoder, x):
    y, J, _ = _net_of(decoder).query(x, order=1)
    return y[:, 0], J[:, 0, :]


def _value(decoder, x):
    y, _, _ = _net_of(decoder).query(x, order=0)
    return y[:, 0]


def _levels(level_decoder, xyz_slice):
    """utils.level_querying_with_confidence (utils.py:269-303): arg-max class of the level MLP."""
    net = getattr(level_decoder, "model", level_decoder)
    return net(xyz_slice).argmax(dim=-1)


@torch.no_grad()
def tvsdf_collision_forward(nozzle_pcd, sdf_decoder, slice_decoder, level_decoder, num_classes, current_level_maxslices,
                            current_levels, base_th):
    """level_collision.py:73-237.  -> (coll_depths [N,M] >= 0, coll_mask [N,M] bool, coll_grads [N,M,3])."""
    N, M, _ = nozzle_pcd.shape
    dev = nozzle_pcd.device
    coords = nozzle_pcd.detach().reshape(-1, 3).to(torch.float32).contiguous()
    P = coords.shape[0]
    waypt_levels = current_levels.to(dev).long().repeat_interleave(M, dim=0)                       # :84
    level_maxslices = current_level_maxslices.to(dev).float().repeat_interleave(M, dim=0)          # :85
    coll_depths = torch.zeros(P, device=dev)
    coll_grads = torch.zeros_like(coords)
    coll_mask = torch.zeros(P, dtype=torch.bool, device=dev)

    def finish():
        return F.relu(-coll_depths).view(N, M), coll_mask.view(N, M), coll_grads.view(N, M, 3)

    cube_mask = ((coords >= -1.0) & (coords <= 1.0)).all(dim=-1)                                   # :92
    cube_idx = cube_mask.nonzero(as_tuple=True)[0]
    if cube_idx.numel() == 0:
        return finish()
    sdfs = torch.ones(P, device=dev)                                                               # :100
    sdf_grads = torch.zeros_like(coords)
    cs, cg = _value_grad(sdf_decoder, coords[cube_idx])
    sdfs[cube_idx] = cs
    sdf_grads[cube_idx] = cg

    base = coords[:, 2] < base_th                                                                  # :113-117
    coll_depths[base] = -0.1
    coll_grads[base] = torch.tensor([0.0, 0.0, 1.0], device=dev)
    coll_mask |= base

    inside_idx = (sdfs < 0.0).nonzero(as_tuple=True)[0]                                            # :120
    if inside_idx.numel() == 0:
        return finish()
    in_coords = coords[inside_idx]
    in_wlevels = waypt_levels[inside_idx]
    in_maxsl = level_maxslices[inside_idx]
    in_sdfs, in_sdf_grads = sdfs[inside_idx], sdf_grads[inside_idx]
    in_wmax = in_maxsl.gather(1, in_wlevels.unsqueeze(-1)).squeeze(-1)
    in_slice, in_slice_grads = _value_grad(slice_decoder, in_coords)                               # :137
    in_levels = _levels(level_decoder, torch.cat((in_coords, in_slice.unsqueeze(-1)), dim=1))      # :139-140
    in_coll = (in_levels < in_wlevels) | ((in_levels == in_wlevels) & (in_slice < in_wmax))        # :143-146
    sel = in_coll.nonzero(as_tuple=True)[0]
    if sel.numel() == 0:
        return finish()
    c_coords, c_levels = in_coords[sel], in_levels[sel]
    c_sdfs, c_sdf_grads = in_sdfs[sel], in_sdf_grads[sel]
    c_slice, c_slice_grads = in_slice[sel], in_slice_grads[sel]
    c_maxsl, c_wlevels = in_maxsl[sel], in_wlevels[sel]
    c_wmax = c_maxsl.gather(1, c_wlevels.unsqueeze(-1)).squeeze(-1)
    c_max_single = c_maxsl.gather(1, c_levels.unsqueeze(-1)).squeeze(-1)                           # :166
    c_dist = (c_slice - c_max_single) / torch.norm(c_slice_grads, dim=-1)                          # :169 (Eq. 27)
    c_tvsdf = torch.max(c_sdfs, c_dist)
    tgt = inside_idx[sel]                                                                          # rows of the flat arrays
    final_tvsdf, final_grads = c_tvsdf.clone(), c_sdf_grads.clone()

    re = (c_dist > c_sdfs).nonzero(as_tuple=True)[0]                                               # :173
    if re.numel() > 0:
        r_coords, r_dist, r_slice_grads = c_coords[re], c_dist[re], c_slice_grads[re]
        push1 = r_coords + torch.abs(r_dist).unsqueeze(-1) * F.normalize(r_slice_grads, dim=-1)    # :202
        push1 = torch.clamp(push1, min=-1.0, max=1.0)
        _, g_push1 = _value_grad(slice_decoder, push1)                                             # :205
        push = push1 + torch.max(torch.abs(r_dist).unsqueeze(-1) * 0.5, torch.tensor(0.04, device=dev)) * \
            F.normalize(g_push1, dim=-1)                                                           # :206-207
        push = torch.clamp(push, min=-1.0, max=1.0)
        p_slice = _value(slice_decoder, push)                                                      # :211
        p_levels = _levels(level_decoder, torch.cat((push, p_slice.unsqueeze(-1)), dim=1))
        r_wlevels, r_wmax = c_wlevels[re], c_wmax[re]
        alter = (p_levels < r_wlevels) | ((p_slice < r_wmax) & (p_levels == r_wlevels))            # :216-217
        final_tvsdf[re] = torch.where(alter, c_sdfs[re], r_dist)                                   # :218-219
        final_grads[re] = torch.where(alter.unsqueeze(-1), c_sdf_grads[re], r_slice_grads)
    coll_depths[tgt] = final_tvsdf                                                                 # :227-231
    coll_grads[tgt] = final_grads
    coll_mask[tgt] = True
    return finish()


class CollTVSDF(torch.autograd.Function):
    """Drop-in for level_collision.CollTVSDF: same argument order, same outputs, same backward (:239-255)."""

    @staticmethod
    def forward(ctx, nozzle_pcd, sdf_decoder, slice_decoder, level_decoder, num_classes, current_level_maxslices,
                current_levels, base_th):
        if not nozzle_pcd.is_cuda:
            raise RuntimeError("inf3dp: CollTVSDF needs CUDA tensors (no CPU fallback)")
        depths, mask, grads = tvsdf_collision_forward(nozzle_pcd, sdf_decoder, slice_decoder, level_decoder, num_classes,
                                                      current_level_maxslices, current_levels, base_th)
        ctx.save_for_backward(grads)
        ctx.mark_non_differentiable(mask)
        return depths, mask

    @staticmethod
    def backward(ctx, d_depths, d_mask=None):
        (grads,) = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return (None,) * 8
        N = grads.shape[0]
        g = -F.normalize(grads, dim=-1, eps=1e-8)
        return g * d_depths.view(N, -1, 1), None, None, None, None, None, None, None


def quaternion_rotate(q, v):
    """utils.py:587-594: rotate v by the quaternion q = (w, x, y, z), broadcasting."""
    q_vec, q_scalar = q[..., 1:], q[..., :1]
    t = 2 * torch.cross(q_vec.expand(torch.broadcast_shapes(q_vec.shape, v.shape)), v.expand(torch.broadcast_shapes(q_vec.shape, v.shape)), dim=-1)
    return v + q_scalar * t + torch.cross(q_vec.expand(t.shape), t, dim=-1)


@torch.no_grad()
def transform_nozzle_pcd_with_quaternion(nozzle_pcd, waypts, quaternions):
    """utils.py:612-621: nozzle point cloud [M,3] posed at every waypoint [N,3] by its quaternion [N,4] -> [N,M,3]."""
    q = F.normalize(quaternions, dim=-1)
    return quaternion_rotate(q[:, None, :], nozzle_pcd[None, :, :]) + waypts[:, None, :]


@torch.no_grad()
def quat_collision_sweep(waypts, waypt_levels, current_level_maxslice, nozzle_pcd, quat_decoder, sdf_decoder, slice_decoder,
                         level_decoder, num_classes, base_th, waypoint_chunk=1 << 18):
    """The collision query of level_collision.quat_collision_test (:451-...) for many waypoints, device resident:
    quaternion field at the waypoints (utils.quat_field_querying) -> posed nozzle clouds -> TV-SDF collision evaluation,
    in chunks of waypoints (the unit the path shards by).  -> (depth per waypoint = max over its nozzle points [N],
    waypoint collision mask [N])."""
    n = waypts.shape[0]
    depth = torch.empty(n, device=waypts.device)
    hit = torch.empty(n, dtype=torch.bool, device=waypts.device)
    qnet = _net_of(quat_decoder)
    for h in range(0, n, waypoint_chunk):
        w = waypts[h:h + waypoint_chunk]
        q, _, _ = qnet.query(w, order=0)
        pcd = transform_nozzle_pcd_with_quaternion(nozzle_pcd, w, q)
        d, m, _ = tvsdf_collision_forward(pcd, sdf_decoder, slice_decoder, level_decoder, num_classes,
                                          current_level_maxslice[h:h + waypoint_chunk], waypt_levels[h:h + waypoint_chunk], base_th)
        depth[h:h + waypoint_chunk] = d.max(dim=-1).values
        hit[h:h + waypoint_chunk] = m.any(dim=-1)
    return depth, hit


def collision_evaluation(nozzle_pcd, sdf_decoder, slice_decoder, level_decoder, num_classes, current_level_maxslice,
                         waypt_levels, base_th):
    """level_collision.py:259-281 -> (coll_depths [N,M], coll_mask [N,M], waypt_coll_mask [N])."""
    depths, mask = CollTVSDF.apply(nozzle_pcd, sdf_decoder, slice_decoder, level_decoder, num_classes,
                                   current_level_maxslice, waypt_levels, base_th)
    return depths, mask, mask.any(dim=-1)
