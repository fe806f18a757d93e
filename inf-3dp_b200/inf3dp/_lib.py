"""ctypes binding of libinf3dp.so (the C ABI declared in include/inf3dp.h).

The library is built in-tree by inf-3dp_b200/csrc/build.sh (nvcc, sm_100a).  There is no fallback: if the shared
object is missing or does not load, importing any compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libinf3dp.so")

OK, EINVAL, ECUDA, EWORKSPACE, EOVERFLOW = 0, 1, 2, 3, 4
ACT_SINE, ACT_RELU = 0, 1
MODE_AUTO, MODE_SIMT, MODE_TC_PRECISE, MODE_TC_FAST, MODE_TC_REVERSE = 0, 1, 2, 3, 4
MODES = {"auto": MODE_AUTO, "simt": MODE_SIMT, "tc_precise": MODE_TC_PRECISE, "tc_fast": MODE_TC_FAST,
         "tc_reverse": MODE_TC_REVERSE}

_vp, _i, _i64, _sz, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float

# name -> (restype, argtypes): mirrors include/inf3dp.h one to one
SIGNATURES = {
    "inf3dp_last_error": (ctypes.c_char_p, []),
    "inf3dp_version": (_i, []),
    "inf3dp_siren_release": (_i, [_vp]),
    "inf3dp_siren_packed_bytes": (_sz, [_i, _i, _i, _i]),
    "inf3dp_siren_pack_weights": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _sz, _vp]),
    "inf3dp_siren_jet": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _vp]),
    "inf3dp_siren_jet_workspace_bytes": (_sz, []),
    "inf3dp_siren_set_sm_limit": (_i, [_i]),
    "inf3dp_siren_jet_ws": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "inf3dp_principal_curvatures": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "inf3dp_unit_normals": (_i, [_vp, _i64, _f, _vp, _vp]),
    "inf3dp_unit_normals_rows": (_i, [_vp, _i64, _f, _vp, _vp]),
    "inf3dp_siren_value_grad_rows": (_i, [_vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "inf3dp_siren_value_grad_rows_peers": (_i, [_vp, _vp, _i64, _vp, _i, _vp, _vp, _sz, _vp]),
    "inf3dp_extract_isocontours_workspace_bytes": (_sz, [_i, _i, _i]),
    "inf3dp_extract_isocontours_workspace_bytes_segments": (_sz, [_i, _i, _i, _i64]),
    "inf3dp_extract_isocontours": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "inf3dp_extract_isocontours_status": (_i, [_vp, _i, _vp, _vp]),
    "inf3dp_extract_isocontours_required_segments": (_i64, [_vp, _i, _vp]),
    "inf3dp_extract_isocontours_compact_rows": (_i64, [_i, _i, _i, _sz]),
    "inf3dp_extract_isocontours_compact": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "inf3dp_fields_intersection_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "inf3dp_fields_intersection": (_i, [_vp] * 7 + [_i] * 6 + [_vp, _vp, _vp, _vp, _sz, _vp]),
    "inf3dp_fields_intersection_grid": (_i, [_vp] * 6 + [_i] * 6 + [_vp, _vp, _vp, _vp, _sz, _vp]),
    "inf3dp_coll_pose": (_i, [_vp, _vp, _vp, _i64, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "inf3dp_coll_inside": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "inf3dp_coll_pack4": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "inf3dp_coll_predicate": (_i, [_vp, _vp, _i, _i64, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "inf3dp_coll_push": (_i, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "inf3dp_coll_alter": (_i, [_vp, _vp, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "inf3dp_coll_gather_x": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "inf3dp_coll_scatter_grads": (_i, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "inf3dp_grid_queries": (_i, [_i, _f, _i64, _i64, _vp, _vp]),
    "inf3dp_marching_cubes_workspace_bytes": (_sz, [_i, _i, _i]),
    "inf3dp_marching_cubes_count": (_i, [_vp, _i, _i, _i, _f, _vp, _sz, _vp, _vp]),
    "inf3dp_marching_cubes_emit": (_i, [_vp, _i, _i, _i, _f, _f, _f, _f, _i, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _sz, _vp]),
}

_lib = None


class Inf3dpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"inf3dp error {code}: {msg}")
        self.code = code


def lib():
    """Load libinf3dp.so (once).  Fails loudly: there is no Python/CPU fallback for any compute path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with inf-3dp_b200/csrc/build.sh "
                              "(or __graft_entry__.build()); inf3dp has no CPU or pure-PyTorch fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != OK:
        raise Inf3dpError(rc, lib().inf3dp_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
