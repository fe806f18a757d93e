"""Synthetic inputs and row-stream helpers shared by bench.py and tests/ (cfg 3 and the a6 size of
slice_toolpath.py:125).  Plain numpy; nothing here touches the GPU or the oracle."""
from __future__ import annotations

import numpy as np

def torus_mesh(nu, nv, R=0.6, r=0.3, jitter=1e-4, seed=0):
    """Closed UV-torus, nu*nv quads split into 2 triangles each -> V [nu*nv,3] f32, F [2*nu*nv,3] i32."""
    rng = np.random.default_rng(seed)
    u = np.arange(nu) * (2 * np.pi / nu)
    v = np.arange(nv) * (2 * np.pi / nv)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    x = (R + r * np.cos(vv)) * np.cos(uu)
    y = (R + r * np.cos(vv)) * np.sin(uu)
    z = r * np.sin(vv)
    V = np.stack([x, y, z], -1).reshape(-1, 3)
    V = V + rng.normal(0.0, jitter, V.shape)
    i = np.arange(nu)[:, None]
    j = np.arange(nv)[None, :]
    a = (i * nv + j).ravel()
    b = (((i + 1) % nu) * nv + j).ravel()
    c = (((i + 1) % nu) * nv + (j + 1) % nv).ravel()
    d = (i * nv + (j + 1) % nv).ravel()
    F = np.concatenate([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], 0)
    # interleave the two triangles of each quad so neighbouring faces stay close in index
    F = F.reshape(2, -1, 3).transpose(1, 0, 2).reshape(-1, 3)
    return V.astype(np.float32), F.astype(np.int32)


def split_rows(merged_k, num):
    """Split one isovalue's row stream at the -1e9 separator rows (toolpath_utils.py:65-82) -> list of [n,3]."""
    sep = np.where(merged_k[:, 0] < -1e2)[0]
    out, start = [], 0
    for s in sep[:num] if num is not None else sep:
        out.append(merged_k[start:s])
        start = s + 1
    return out


def rows_used(merged_k, num):
    """S_k + C_k recovered from the dense stream: index of the last separator + 1."""
    if num == 0:
        return 0
    sep = np.where(merged_k[:, 0] < -1e2)[0]
    return int(sep[num - 1]) + 1


CORNER_OFFSETS = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]


def smooth_grids(N, poly=False):
    """Three smooth [N,N,N] float64 sample grids on [-1,1]^3 (slice-like ~z, lattice-like ~x and ~y).
    poly=True uses +, -, * only (bit-reproducible on any host: what committed goldens are recorded on)."""
    g = np.stack(np.meshgrid(*[np.linspace(-1, 1, N)] * 3, indexing="ij"), -1)
    x = g[..., 0]
    wave = 0.1 * (3.0 * x - 4.5 * x * x * x) if poly else 0.1 * np.sin(3 * x)
    return [g[..., 2] + wave, g[..., 0] + 0.2 * g[..., 1] * g[..., 1], g[..., 1] - 0.15 * g[..., 2] * g[..., 0]]


def corner_expanded_fields(D, seed=0, nan_frac=0.2, poly=False):
    """The inputs of get_fields_intersection_inter_slice as the reference's host code builds them
    (toolpath_utils.py:997-1033): three corner-expanded [D,D,D,8] float32 tensors (corner c = offsets CORNER_OFFSETS[c])
    and the [D,D,D,3] voxel centres; `nan_frac` of the voxels NaN-masked (outside the SDF, toolpath_utils.py:1071-1084)."""
    rng = np.random.default_rng(seed)
    base = smooth_grids(D + 1, poly)
    out = [np.ascontiguousarray(np.stack([b[o[0]:o[0] + D, o[1]:o[1] + D, o[2]:o[2] + D] for o in CORNER_OFFSETS], -1).astype(np.float32))
           for b in base]
    idx = np.stack(np.meshgrid(*[np.arange(D)] * 3, indexing="ij"), -1).astype(np.float32)
    centers = ((idx + 0.5) * (2.0 / D) - 1.0).astype(np.float32)
    if nan_frac > 0:
        nanmask = rng.random((D, D, D)) < nan_frac
        for c in out:
            c[nanmask] = np.nan
    return out, centers
