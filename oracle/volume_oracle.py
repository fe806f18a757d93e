"""ORACLE (test infrastructure, not product code) -- ctypes front end of oracle/volume_oracle.c.

Builds oracle/_build/libvolume_oracle.so with gcc on first use (or via __graft_entry__.build()).
See volume_oracle.c for the reference file:line each function follows and for the pinning statement
(grid queries / corner expansion / voxel centres: pinned to the reference; marching cubes: PARITY UNPINNED).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "volume_oracle.c")
_TABLE = os.path.join(os.path.dirname(_HERE), "inf-3dp_b200", "csrc", "mc_table.inc")
_OUT = os.path.join(_HERE, "_build", "libvolume_oracle.so")
_lib = None


def build(force=False):
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    newest = max(os.path.getmtime(_SRC), os.path.getmtime(_TABLE))
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < newest:
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               "-o", _OUT, _SRC, "-lm"])
    return _OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def grid_queries(N, cube_size=1.0, start=0, count=None):
    """sdf_meshing.py:46-65 -> samples [count,3] f32 of the index range [start, start+count)."""
    count = N ** 3 - start if count is None else count
    out = np.empty((count, 3), dtype=np.float32)
    lib().oracle_grid_queries(ctypes.c_int(N), ctypes.c_float(cube_size), ctypes.c_longlong(start),
                              ctypes.c_longlong(count), _p(out, ctypes.c_float))
    return out


def corner_fields(grid):
    """toolpath_utils.py:986-1019 -> [Gx-1,Gy-1,Gz-1,8] f32."""
    g = np.ascontiguousarray(grid, dtype=np.float32)
    Gx, Gy, Gz = g.shape
    out = np.empty((Gx - 1, Gy - 1, Gz - 1, 8), dtype=np.float32)
    lib().oracle_corner_fields(_p(g, ctypes.c_float), Gx, Gy, Gz, _p(out, ctypes.c_float))
    return out


def voxel_centers(N):
    """toolpath_utils.py:1021-1033 -> [N-1,N-1,N-1,3] f32."""
    out = np.empty((N - 1, N - 1, N - 1, 3), dtype=np.float32)
    lib().oracle_voxel_centers(N, _p(out, ctypes.c_float))
    return out


def marching_cubes(volume, level=0.0, spacing=(1.0, 1.0, 1.0), gradient_direction="descent"):
    """-> verts [nv,3] f32, faces [nf,3] i32, normals [nv,3] f32, values [nv] f32 (skimage's return order)."""
    v = np.ascontiguousarray(volume, dtype=np.float32)
    Gx, Gy, Gz = v.shape
    sp = np.asarray(spacing, dtype=np.float32)
    descent = {"descent": 1, "ascent": 0}[gradient_direction]
    counts = np.zeros(2, dtype=np.int64)
    f = lib().oracle_marching_cubes
    rc = f(_p(v, ctypes.c_float), Gx, Gy, Gz, ctypes.c_float(level), _p(sp, ctypes.c_float), descent,
           None, None, None, None, _p(counts, ctypes.c_longlong))
    assert rc == 0, rc
    nv, nf = int(counts[0]), int(counts[1])
    verts = np.empty((nv, 3), dtype=np.float32)
    faces = np.empty((nf, 3), dtype=np.int32)
    normals = np.empty((nv, 3), dtype=np.float32)
    values = np.empty(nv, dtype=np.float32)
    rc = f(_p(v, ctypes.c_float), Gx, Gy, Gz, ctypes.c_float(level), _p(sp, ctypes.c_float), descent,
           _p(verts, ctypes.c_float), _p(faces, ctypes.c_int32), _p(normals, ctypes.c_float), _p(values, ctypes.c_float),
           _p(counts, ctypes.c_longlong))
    if rc != 0:
        raise RuntimeError(f"oracle_marching_cubes failed with code {rc}")
    assert (int(counts[0]), int(counts[1])) == (nv, nf)
    return verts, faces, normals, values


# ---- mesh properties used by the tests (numpy, independent of both implementations) ---------------------------------
def mesh_edge_stats(faces):
    """-> (number of undirected edges, edges not shared by exactly two triangles, directed edges used more than once)."""
    f = np.asarray(faces, dtype=np.int64)
    d = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
    key_d = d[:, 0] * (f.max() + 1 if len(f) else 1) + d[:, 1]
    _, cd = np.unique(key_d, return_counts=True)
    u = np.sort(d, axis=1)
    key_u = u[:, 0] * (f.max() + 1 if len(f) else 1) + u[:, 1]
    _, cu = np.unique(key_u, return_counts=True)
    return len(cu), int((cu != 2).sum()), int((cd > 1).sum())


def euler_characteristic(n_verts, faces):
    e, _, _ = mesh_edge_stats(faces)
    return n_verts - e + len(faces)
