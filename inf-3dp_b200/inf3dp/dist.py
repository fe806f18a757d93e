"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink), replicated weights, points or
isovalues sharded contiguously, one final all-gather.  The reference has no distributed code.
Works with the gloo backend on CPU tensors for host-side tests of the sharding logic.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous ceil(n/world) ranges; the last ranks may be short or empty."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


def all_gather_rows(local, n_total, group=None):
    """Gather per-rank row slabs (each rank holds rows shard_range(n_total, rank, world)) into [n_total, ...]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    lo, hi, per = shard_range(n_total, rank, world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:hi - lo] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return out[:n_total]


def sharded_value_grad(net, coords, mode=None, group=None, gather=True):
    """Each rank evaluates its contiguous point range; optionally all-gathers [N,4] = (value, grad)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = coords.shape[0]
    lo, hi, _ = shard_range(n, rank, world)
    y, J, _ = net.query(coords[lo:hi], order=1, mode=mode)
    local = torch.cat([y[:, :1], J[:, 0, :]], dim=1)
    return all_gather_rows(local, n, group) if gather else local


def sharded_isocontours(vertices, faces, fields, iso_values, group=None, gather=True):
    """Isovalues are sharded (chaining needs all segments of one isovalue on one rank); mesh and field replicated."""
    from .contours import extract_isocontours_raw
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    K = iso_values.numel()
    lo, hi, _ = shard_range(K, rank, world)
    merged, nums = extract_isocontours_raw(vertices, faces, fields, iso_values[lo:hi].contiguous())
    if not gather or world == 1:
        return merged, nums
    return all_gather_rows(merged, K, group), all_gather_rows(nums, K, group)
