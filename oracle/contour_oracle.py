"""ORACLE (test infrastructure, not product code) -- ctypes front end of oracle/contour_oracle.c.

Builds oracle/_build/libcontour_oracle.so with gcc on first use (or via __graft_entry__.build()).
See contour_oracle.c for the reference file:line each function follows and for the pinning statement.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "contour_oracle.c")
_OUT = os.path.join(_HERE, "_build", "libcontour_oracle.so")
_lib = None


def build(force=False):
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               "-o", _OUT, _SRC, "-lm"])
    return _OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_intersection_unmatched.restype = ctypes.c_long
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def extract_isocontours(vertices, faces, fields, iso_values):
    """-> merged [K,F,3] f32, contour_nums [K] i32, seg_counts [K] i32 (the reference's pairpoint_nums)."""
    V = np.ascontiguousarray(vertices, dtype=np.float32)
    F = np.ascontiguousarray(faces, dtype=np.int32)
    f = np.ascontiguousarray(fields, dtype=np.float32)
    iso = np.ascontiguousarray(iso_values, dtype=np.float32)
    K, nF = len(iso), len(F)
    merged = np.empty((K, nF, 3), dtype=np.float32)
    nums = np.zeros(K, dtype=np.int32)
    segs = np.zeros(K, dtype=np.int32)
    rc = lib().oracle_extract_isocontours(_p(V, ctypes.c_float), _p(F, ctypes.c_int32), _p(f, ctypes.c_float),
                                          _p(iso, ctypes.c_float), len(V), nF, K, _p(merged, ctypes.c_float),
                                          _p(nums, ctypes.c_int32), _p(segs, ctypes.c_int32))
    if rc != 0:
        raise MemoryError("oracle_extract_isocontours")
    return merged, nums, segs


def _prep_fields(slice_f, l1_f, l2_f, s_iso, l1_iso, l2_iso, centers):
    s = np.ascontiguousarray(slice_f, dtype=np.float32)
    a = np.ascontiguousarray(l1_f, dtype=np.float32)
    b = np.ascontiguousarray(l2_f, dtype=np.float32)
    si = np.ascontiguousarray(s_iso, dtype=np.float32)
    ai = np.ascontiguousarray(l1_iso, dtype=np.float32)
    bi = np.ascontiguousarray(l2_iso, dtype=np.float32)
    c = np.ascontiguousarray(centers, dtype=np.float32)
    assert s.shape == a.shape == b.shape and s.shape[-1] == 8 and c.shape == s.shape[:3] + (3,)
    return s, a, b, si, ai, bi, c


def fields_intersection(slice_f, l1_f, l2_f, s_iso, l1_iso, l2_iso, centers, winner_rule=0):
    """-> mask [Ns,N1,N2] bool, idx [Ns,N1,N2,3] i16, pts [Ns,N1,N2,3] f32."""
    s, a, b, si, ai, bi, c = _prep_fields(slice_f, l1_f, l2_f, s_iso, l1_iso, l2_iso, centers)
    Dx, Dy, Dz = s.shape[:3]
    Ns, N1, N2 = len(si), len(ai), len(bi)
    mask = np.zeros((Ns, N1, N2), dtype=np.uint8)
    idx = np.zeros((Ns, N1, N2, 3), dtype=np.int16)
    pts = np.zeros((Ns, N1, N2, 3), dtype=np.float32)
    lib().oracle_fields_intersection(_p(s, ctypes.c_float), _p(a, ctypes.c_float), _p(b, ctypes.c_float),
                                     _p(si, ctypes.c_float), _p(ai, ctypes.c_float), _p(bi, ctypes.c_float),
                                     _p(c, ctypes.c_float), Dx, Dy, Dz, Ns, N1, N2, _p(mask, ctypes.c_uint8),
                                     _p(idx, ctypes.c_int16), _p(pts, ctypes.c_float), int(winner_rule))
    return mask.astype(bool), idx, pts


def intersection_unmatched(slice_f, l1_f, l2_f, s_iso, l1_iso, l2_iso, centers, mask, pts, tol=1e-5):
    """Number of cells of (mask, pts) that no schedule of the reference kernel could have produced."""
    s, a, b, si, ai, bi, c = _prep_fields(slice_f, l1_f, l2_f, s_iso, l1_iso, l2_iso, centers)
    Dx, Dy, Dz = s.shape[:3]
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    p = np.ascontiguousarray(pts, dtype=np.float32)
    return int(lib().oracle_intersection_unmatched(
        _p(s, ctypes.c_float), _p(a, ctypes.c_float), _p(b, ctypes.c_float), _p(si, ctypes.c_float),
        _p(ai, ctypes.c_float), _p(bi, ctypes.c_float), _p(c, ctypes.c_float), Dx, Dy, Dz, len(si), len(ai),
        len(bi), _p(m, ctypes.c_uint8), _p(p, ctypes.c_float), ctypes.c_float(tol)))


# ---- synthetic inputs / row-stream helpers: they live in the product package's test-support module (bench.py's product
# arm uses them too and must not import oracle/); re-exported here for the tests that already use co.torus_mesh etc.
import sys as _sys
_PKG = os.path.join(os.path.dirname(_HERE), "inf-3dp_b200")
if _PKG not in _sys.path:
    _sys.path.insert(0, _PKG)
from inf3dp.testing import torus_mesh, split_rows, rows_used  # noqa: E402,F401
