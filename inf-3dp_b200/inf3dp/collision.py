"""Device-resident TV-SDF collision evaluation of nozzle point clouds against the printed part.

Reference: level_collision.CollTVSDF (level_collision.py:68-255, paper section 5.1) and its wrapper
collision_evaluation (:259-281).  The reference runs the same sequence through utils.field_query_with_grads /
level_querying_with_confidence, which build an autograd graph per chunk and bounce every intermediate through host
tensors (utils.py:232-303), and glues them with boolean masks / nonzero / gather / where.  Here the field sweeps are
launches of the fused SIREN kernels on compact device lists, and every step between two sweeps is ONE fused predicate /
compaction kernel (csrc/collision.cu, include/inf3dp.h: inf3dp_coll_*).  The SDF gradient is
queried lazily, only for the points whose result carries it.  Same predicates, same Eq. 27, same two-step push and the
same backward (negated, normalised saved directions times the upstream gradient).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from .query import _net_of


def _rows(decoder, x):
    """(value, grad) rows [n,4] of a single-output field network: the reverse-mode engine's packed output when the network
    is the flagship configuration, otherwise the generic order-1 query packed with torch."""
    net = _net_of(decoder)
    if (getattr(net, "out_features", 0) == 1 and getattr(net, "in_features", 0) == 3 and getattr(net, "type", "") == "sine"
            and getattr(net, "hidden_features", 0) == 256 and getattr(net, "num_hidden_layers", 0) == 3):
        return net.query_rows(x)
    y, J, _ = net.query(x, order=1)
    return torch.cat([y[:, :1], J[:, 0, :]], dim=1).contiguous()


def _logits(level_decoder, x4):
    net = getattr(level_decoder, "model", level_decoder)
    return net(x4).to(torch.float32).contiguous()


class _Surface:
    """Where finished points report: per point (CollTVSDF surface) and / or per waypoint (sweep)."""

    def __init__(self, dev, P=0, Nw=0, per_point=False, per_waypoint=False):
        z = lambda *shape, dtype=torch.float32: torch.zeros(shape, dtype=dtype, device=dev)
        self.depth_p = z(P) if per_point else None
        self.mask_p = z(P, dtype=torch.uint8) if per_point else None
        self.grads_p = z(P, 3) if per_point else None
        self.depth_w = z(Nw, dtype=torch.int32) if per_waypoint else None      # IEEE bits of the max depth (depths >= 0)
        self.hit_w = z(Nw, dtype=torch.uint8) if per_waypoint else None

    def ptrs(self):
        return [_lib.ptr(t) for t in (self.depth_p, self.mask_p, self.grads_p, self.depth_w, self.hit_w)]


@torch.no_grad()
def _fused_collision(surface, waypts, quats, cloud, Nw, M, sdf_decoder, slice_decoder, level_decoder, num_classes,
                     current_level_maxslices, current_levels, base_th, want_grads):
    """level_collision.py:73-237 as six field sweeps glued by the fused predicate / compaction stages of csrc/collision.cu
    (include/inf3dp.h: inf3dp_coll_*).  `cloud` is the nozzle cloud [M,3] when `quats` [Nw,4] pose it at `waypts` [Nw,3]
    (the posed cloud [Nw,M,3] is never materialised), or the posed cloud [Nw*M,3] itself when quats is None."""
    dev = cloud.device
    L = _lib.lib()
    P = int(Nw) * int(M)
    C = int(num_classes)
    st = lambda: _lib.stream_ptr(dev)
    e = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype, device=dev)
    levels_w = current_levels.to(device=dev, dtype=torch.int32).contiguous()
    maxsl = current_level_maxslices.to(device=dev, dtype=torch.float32).contiguous()
    if maxsl.dim() != 2 or maxsl.shape[1] != C or maxsl.shape[0] != Nw or levels_w.numel() != Nw:
        raise RuntimeError("collision: current_level_maxslices must be [N, num_classes] and current_levels [N]")
    if Nw > 0:
        lo_l, hi_l = torch.aminmax(levels_w)
        if int(lo_l) < 0 or int(hi_l) >= C:      # the kernels index maxslices[w, level]: the reference's gather would assert
            raise RuntimeError(f"collision: current_levels must lie in [0, {C}) (got {int(lo_l)} .. {int(hi_l)})")
    counters = e(8, dtype=torch.int32)
    out = surface.ptrs()
    with torch.cuda.device(dev):
        cube_x, cube_id = e(max(P, 1), 3), e(max(P, 1), dtype=torch.int32)
        _lib.check(L.inf3dp_coll_pose(_lib.ptr(waypts), _lib.ptr(quats), _lib.ptr(cloud), Nw, M, float(base_th), _lib.ptr(cube_x),
                                      _lib.ptr(cube_id), _lib.ptr(counters), *out, st()))
        n_cube = int(counters[0].item())
        if n_cube == 0:
            return
        # SDF value only: its gradient is needed for the few points whose result carries it, queried lazily below (:100-107)
        sdf_v, _, _ = _net_of(sdf_decoder).query(cube_x[:n_cube], order=0)
        in_x, in_id, in_sdf = e(n_cube, 3), e(n_cube, dtype=torch.int32), e(n_cube)
        _lib.check(L.inf3dp_coll_inside(_lib.ptr(sdf_v), n_cube, _lib.ptr(cube_x), _lib.ptr(cube_id), _lib.ptr(in_x),
                                        _lib.ptr(in_id), _lib.ptr(in_sdf), _lib.ptr(counters), M, *out, st()))
        del cube_x, cube_id, sdf_v
        n_in = int(counters[1].item())
        if n_in == 0:
            return
        rows = _rows(slice_decoder, in_x[:n_in])                                                        # :137
        x4 = e(n_in, 4)
        _lib.check(L.inf3dp_coll_pack4(_lib.ptr(in_x), _lib.ptr(rows), 4, n_in, _lib.ptr(x4), st()))
        logits = _logits(level_decoder, x4)                                                             # :139-142
        if logits.shape[1] != C:
            raise RuntimeError(f"collision: the level classifier has {logits.shape[1]} classes, num_classes is {C}")
        del x4
        re_x, re_src, re_dist = e(n_in, 3), e(n_in, dtype=torch.int32), e(n_in)
        grad_src = e(n_in, dtype=torch.int32) if want_grads else None
        _lib.check(L.inf3dp_coll_predicate(_lib.ptr(rows), _lib.ptr(logits), C, n_in, _lib.ptr(in_x), _lib.ptr(in_id),
                                           _lib.ptr(in_sdf), _lib.ptr(levels_w), _lib.ptr(maxsl), M, _lib.ptr(re_x),
                                           _lib.ptr(re_src), _lib.ptr(re_dist), _lib.ptr(grad_src), int(want_grads),
                                           _lib.ptr(counters), *out, st()))
        del logits
        n_re = int(counters[2].item())
        if n_re > 0:
            rows2 = _rows(slice_decoder, re_x[:n_re])                                                   # :205
            push_x = e(n_re, 3)
            _lib.check(L.inf3dp_coll_push(_lib.ptr(rows2), n_re, _lib.ptr(re_x), _lib.ptr(re_dist), _lib.ptr(push_x), st()))
            del rows2
            y2, _, _ = _net_of(slice_decoder).query(push_x, order=0)                                    # :211
            x4 = e(n_re, 4)
            _lib.check(L.inf3dp_coll_pack4(_lib.ptr(push_x), _lib.ptr(y2), 1, n_re, _lib.ptr(x4), st()))
            logits2 = _logits(level_decoder, x4)                                                        # :212-213
            _lib.check(L.inf3dp_coll_alter(_lib.ptr(y2), _lib.ptr(logits2), C, n_re, _lib.ptr(re_src), _lib.ptr(re_dist),
                                           _lib.ptr(rows), _lib.ptr(in_id), _lib.ptr(in_sdf), _lib.ptr(levels_w),
                                           _lib.ptr(maxsl), M, _lib.ptr(grad_src), int(want_grads), _lib.ptr(counters), *out, st()))
        if want_grads:
            n_g = int(counters[3].item())
            if n_g > 0:
                gx = e(n_g, 3)
                _lib.check(L.inf3dp_coll_gather_x(_lib.ptr(grad_src), n_g, _lib.ptr(in_x), _lib.ptr(gx), st()))
                rows3 = _rows(sdf_decoder, gx)
                _lib.check(L.inf3dp_coll_scatter_grads(_lib.ptr(rows3), n_g, _lib.ptr(grad_src), _lib.ptr(in_id),
                                                       _lib.ptr(surface.grads_p), st()))


@torch.no_grad()
def tvsdf_collision_forward(nozzle_pcd, sdf_decoder, slice_decoder, level_decoder, num_classes, current_level_maxslices,
                            current_levels, base_th):
    """level_collision.py:73-237.  -> (coll_depths [N,M] >= 0, coll_mask [N,M] bool, coll_grads [N,M,3])."""
    if not nozzle_pcd.is_cuda:
        raise RuntimeError("inf3dp: the collision query needs CUDA tensors (no CPU fallback)")
    N, M, _ = nozzle_pcd.shape
    dev = nozzle_pcd.device
    coords = nozzle_pcd.detach().reshape(-1, 3).to(torch.float32).contiguous()
    if N * M >= 1 << 30:
        raise RuntimeError("inf3dp: CollTVSDF handles < 2^30 points per call; chunk the waypoints (quat_collision_sweep does)")
    surf = _Surface(dev, P=N * M, per_point=True)
    _fused_collision(surf, None, None, coords, N, M, sdf_decoder, slice_decoder, level_decoder, num_classes,
                     current_level_maxslices, current_levels, base_th, want_grads=True)
    return surf.depth_p.view(N, M), surf.mask_p.view(N, M).bool(), surf.grads_p.view(N, M, 3)


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
    quaternion field at the waypoints (utils.quat_field_querying) -> nozzle cloud posed on the fly -> TV-SDF collision
    evaluation, in chunks of waypoints (the unit the path shards by).  Neither the posed clouds [N,M,3] nor any [N,M] array
    is materialised: finished points are reduced per waypoint inside the stage kernels.
    -> (depth per waypoint = max over its nozzle points [N], waypoint collision mask [N])."""
    if not waypts.is_cuda:
        raise RuntimeError("inf3dp: the collision query needs CUDA tensors (no CPU fallback)")
    n = waypts.shape[0]
    dev = waypts.device
    M = nozzle_pcd.shape[0]
    waypoint_chunk = max(1, min(int(waypoint_chunk), ((1 << 30) - 1) // max(M, 1)))
    depth = torch.empty(n, device=dev)
    hit = torch.empty(n, dtype=torch.bool, device=dev)
    qnet = _net_of(quat_decoder)
    nozzle = nozzle_pcd.detach().to(torch.float32).contiguous()
    for h in range(0, n, waypoint_chunk):
        w = waypts[h:h + waypoint_chunk].detach().to(torch.float32).contiguous()
        nw = w.shape[0]
        q, _, _ = qnet.query(w, order=0)
        surf = _Surface(dev, Nw=nw, per_waypoint=True)
        _fused_collision(surf, w, q.contiguous(), nozzle, nw, M, sdf_decoder, slice_decoder, level_decoder, num_classes,
                         current_level_maxslice[h:h + waypoint_chunk], waypt_levels[h:h + waypoint_chunk], base_th,
                         want_grads=False)
        depth[h:h + nw] = surf.depth_w.view(torch.float32)
        hit[h:h + nw] = surf.hit_w.bool()
    return depth, hit


@torch.no_grad()
def quat_collision_sweep_unfused(waypts, waypt_levels, current_level_maxslice, nozzle_pcd, quat_decoder, sdf_decoder,
                                 slice_decoder, level_decoder, num_classes, base_th, waypoint_chunk=1 << 16):
    """The same sweep through the per-point surface (posed clouds materialised, [N,M] results reduced with torch): the
    composition a caller of the reference would write (utils.transform_nozzle_pcd_with_quaternion + CollTVSDF); kept as the
    comparator of quat_collision_sweep in tests/."""
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
