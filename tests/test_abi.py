"""CPU: the C-ABI library loads and exports every symbol include/inf3dp.h declares; argument errors are
reported through return codes + inf3dp_last_error() (no compute is launched here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "inf3dp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(inf3dp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from inf3dp import _lib
    assert os.path.exists(_lib.LIB_PATH), "libinf3dp.so not built (run inf-3dp_b200/csrc/build.sh)"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/inf3dp.h but not exported"
    assert set(_lib.SIGNATURES) == set(names), "ctypes table out of sync with the header"


def test_argument_errors_use_return_codes():
    from inf3dp import _lib
    L = _lib.lib()
    assert L.inf3dp_version() >= 100
    assert L.inf3dp_siren_packed_bytes(3, 256, 3, 1) > 790_000
    assert L.inf3dp_siren_packed_bytes(3, 128, 3, 1) == 0  # unsupported width
    rc = L.inf3dp_siren_jet(None, None, 5, 1, None, None, None, 0, None)
    assert rc == _lib.EINVAL and b"null" in L.inf3dp_last_error()
    rc = L.inf3dp_extract_isocontours(None, None, None, None, 10, 10, -1, None, None, None, 0, None)
    assert rc == _lib.EINVAL
    assert L.inf3dp_extract_isocontours_workspace_bytes(1000, 2000, 10) > 0
    assert L.inf3dp_extract_isocontours_workspace_bytes_segments(1000, 2000, 10, 20000) >= \
        L.inf3dp_extract_isocontours_workspace_bytes(1000, 2000, 10)
    assert L.inf3dp_fields_intersection_workspace_bytes(8, 8, 8, 4, 4, 4) > 0
    with pytest.raises(_lib.Inf3dpError):
        _lib.check(L.inf3dp_principal_curvatures(None, None, -1, None, None, None))
    # round-2 entry points: argument errors come back as codes, nothing is launched
    assert L.inf3dp_siren_release(None) == _lib.OK
    assert L.inf3dp_siren_value_grad_rows_peers(None, None, 5, None, 0, None, None, 0, None) == _lib.EINVAL
    assert L.inf3dp_extract_isocontours_required_segments(None, 4, None) == 0
    assert L.inf3dp_coll_pose(None, None, None, 4, 0, 0.0, None, None, None, None, None, None, None, None, None) == _lib.EINVAL
    assert b"bad sizes" in L.inf3dp_last_error()
    assert L.inf3dp_coll_pose(None, None, None, 1 << 20, 1 << 12, 0.0, None, None, None, None, None, None, None, None, None) == _lib.EINVAL
    assert b"2^30" in L.inf3dp_last_error()
    assert L.inf3dp_coll_inside(None, -1, None, None, None, None, None, None, 1, None, None, None, None, None, None) == _lib.EINVAL
    assert L.inf3dp_coll_predicate(None, None, 8, 5, None, None, None, None, None, 4, None, None, None, None, 0, None,
                                   None, None, None, None, None, None) == _lib.EINVAL
    assert L.inf3dp_coll_pack4(None, None, 1, 0, None, None) == _lib.OK        # n = 0: nothing to do
