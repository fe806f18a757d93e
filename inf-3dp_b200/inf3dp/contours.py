"""Python entry points of the native contour operators (positional signatures of the reference's pybind module).

Reference: contour_cuda/ext.cpp:5-17 (extract_isocontours) and :20-46 (get_fields_intersection_inter_slice), bound
as `qupkg.extract_isocontours` (contour_cuda/setup.py:8-12) and called from toolpath_utils.py:33 and :1121.
Argument checking follows CHECK_INPUT (contour_cuda/includes/utils.h:6-8): every argument must be a contiguous
CUDA tensor, otherwise RuntimeError.  The reference dispatches float32/float64; this build implements float32
(the only dtype its callers pass, toolpath_utils.py:27-30, :1110-1116) and raises on anything else.
Like the reference, both calls return only after the device work has finished.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _check_input(t, name, dtype=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must have dtype {dtype} (got {t.dtype})")


def extract_isocontours_raw(vertices, faces, fields, iso_values, sync=True, want_seg_counts=False):
    """-> contours_merged [K,F,3] f32, contour_nums [K] i32 (, seg_counts [K] i32 on the host).

    sync=False returns without reading the status back: a segment-pool overflow (coarse mesh with a dense isovalue set:
    more than nF + 65536 segments) is then NOT detected and the call yields all-zero contours with contour_nums == 0;
    use it only where the pool is known to suffice (benchmarks), never on user data."""
    _check_input(vertices, "vertices", torch.float32)
    _check_input(faces, "faces", torch.int32)
    _check_input(fields, "fields", torch.float32)
    _check_input(iso_values, "iso_values", torch.float32)
    if vertices.dim() != 2 or vertices.shape[1] != 3 or faces.dim() != 2 or faces.shape[1] != 3:
        raise RuntimeError("vertices must be [V,3] and faces [F,3]")
    if fields.numel() != vertices.shape[0]:
        raise RuntimeError("fields must hold one value per vertex")
    dev = vertices.device
    nV, nF, K = vertices.shape[0], faces.shape[0], iso_values.numel()
    L = _lib.lib()
    merged = torch.empty((K, nF, 3), dtype=torch.float32, device=dev)
    nums = torch.empty((K,), dtype=torch.int32, device=dev)
    if K == 0:
        return (merged, nums, torch.zeros(0, dtype=torch.int32)) if want_seg_counts else (merged, nums)
    ws_bytes = L.inf3dp_extract_isocontours_workspace_bytes(nV, nF, K)
    seg_host = (ctypes.c_int32 * K)() if want_seg_counts else None
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        for attempt in range(2):
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(L.inf3dp_extract_isocontours(_lib.ptr(vertices), _lib.ptr(faces), _lib.ptr(fields),
                                                    _lib.ptr(iso_values), nV, nF, K, _lib.ptr(merged), _lib.ptr(nums),
                                                    _lib.ptr(ws), ws_bytes, st))
            if not sync and not want_seg_counts:
                break
            rc = L.inf3dp_extract_isocontours_status(_lib.ptr(ws), K, seg_host, st)
            if rc == _lib.EOVERFLOW and attempt == 0:
                # the segment pool was too small for this mesh/field: the count pass knows how many segments there are
                need = L.inf3dp_extract_isocontours_required_segments(_lib.ptr(ws), K, st)
                ws_bytes = L.inf3dp_extract_isocontours_workspace_bytes_segments(nV, nF, K, need if need > 0 else int(nF) * int(K))
                continue
            _lib.check(rc)
            break
    if want_seg_counts:
        return merged, nums, torch.tensor(list(seg_host), dtype=torch.int32)
    return merged, nums


def extract_isocontours_compact(vertices, faces, fields, iso_values):
    """Compact row streams: -> rows [R,3] f32 (device), row_offsets [K] i64, row_counts [K] i32,
    contour_nums [K] i32 (all on the device).  rows[row_offsets[k] : row_offsets[k] + row_counts[k]] is bit-identical to the
    used prefix of extract_isocontours(...)[0][k]; the dense zero-padded [K,F,3] tensor is never produced."""
    _check_input(vertices, "vertices", torch.float32)
    _check_input(faces, "faces", torch.int32)
    _check_input(fields, "fields", torch.float32)
    _check_input(iso_values, "iso_values", torch.float32)
    if vertices.dim() != 2 or vertices.shape[1] != 3 or faces.dim() != 2 or faces.shape[1] != 3:
        raise RuntimeError("vertices must be [V,3] and faces [F,3]")
    if fields.numel() != vertices.shape[0]:
        raise RuntimeError("fields must hold one value per vertex")
    dev = vertices.device
    nV, nF, K = vertices.shape[0], faces.shape[0], iso_values.numel()
    L = _lib.lib()
    offs = torch.zeros((K,), dtype=torch.int64, device=dev)
    counts = torch.zeros((K,), dtype=torch.int32, device=dev)
    nums = torch.zeros((K,), dtype=torch.int32, device=dev)
    if K == 0:
        return torch.empty((0, 3), dtype=torch.float32, device=dev), offs, counts, nums
    ws_bytes = L.inf3dp_extract_isocontours_workspace_bytes(nV, nF, K)
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        for attempt in range(2):
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            cap_rows = L.inf3dp_extract_isocontours_compact_rows(nV, nF, K, ws_bytes)
            rows = torch.empty((cap_rows, 3), dtype=torch.float32, device=dev)
            _lib.check(L.inf3dp_extract_isocontours_compact(_lib.ptr(vertices), _lib.ptr(faces), _lib.ptr(fields),
                                                            _lib.ptr(iso_values), nV, nF, K, _lib.ptr(rows), cap_rows,
                                                            _lib.ptr(offs), _lib.ptr(counts), _lib.ptr(nums), _lib.ptr(ws),
                                                            ws_bytes, st))
            rc = L.inf3dp_extract_isocontours_status(_lib.ptr(ws), K, None, st)
            if rc == _lib.EOVERFLOW and attempt == 0:
                need = L.inf3dp_extract_isocontours_required_segments(_lib.ptr(ws), K, st)
                ws_bytes = L.inf3dp_extract_isocontours_workspace_bytes_segments(nV, nF, K, need if need > 0 else int(nF) * int(K))
                continue
            _lib.check(rc)
            break
    return rows, offs, counts, nums


def contour_loops(vertices, faces, fields, iso_values):
    """Per-isovalue list of closed / open polylines as host arrays -- what the reference obtains by copying the dense
    [K,F,3] tensor to the host and scanning it for separator rows (toolpath_utils.py:38-82), here from one transfer of
    the used rows only (tens of MB instead of 12*K*F bytes).  -> list over isovalues of lists of [n_i, 3] float32 arrays."""
    rows, offs, counts, nums = extract_isocontours_compact(vertices, faces, fields, iso_values)
    offs_h, counts_h = offs.cpu().numpy(), counts.cpu().numpy()
    total = int((offs_h + counts_h).max()) if len(offs_h) else 0
    rows_h = rows[:total].cpu().numpy()
    out = []
    for k in range(len(offs_h)):
        r = rows_h[offs_h[k]: offs_h[k] + counts_h[k]]
        sep = (r[:, 0] == -1e9).nonzero()[0] if len(r) else []
        loops, start = [], 0
        for e in sep:
            loops.append(r[start:e])
            start = e + 1
        out.append(loops)
    return out


def extract_isocontours(vertices, faces, fields, iso_values):
    """qupkg.extract_isocontours.extract_isocontours -> [contours_merged, contour_nums] (isocontours.cu:292)."""
    merged, nums = extract_isocontours_raw(vertices, faces, fields, iso_values, sync=True)
    return [merged, nums]


def get_fields_intersection_inter_slice(slice_fields, lattice1_fields, lattice2_fields, slice_iso_values,
                                        lattice1_iso_values, lattice2_iso_values, voxel_centers, sync=True):
    """-> [inter_mask bool [Ns,N1,N2], isovalue_indices int16 [Ns,N1,N2,3], intersections f32 [Ns,N1,N2,3]]
    (intersection.cu:285)."""
    names = ["slice_fields", "lattice1_fields", "lattice2_fields", "slice_iso_values", "lattice1_iso_values",
             "lattice2_iso_values", "voxel_centers"]
    args = [slice_fields, lattice1_fields, lattice2_fields, slice_iso_values, lattice1_iso_values, lattice2_iso_values,
            voxel_centers]
    for t, n in zip(args, names):
        _check_input(t, n, torch.float32)
    if slice_fields.dim() != 4 or slice_fields.shape[-1] != 8:
        raise RuntimeError("field tensors must be [Dx,Dy,Dz,8]")
    if lattice1_fields.shape != slice_fields.shape or lattice2_fields.shape != slice_fields.shape:
        raise RuntimeError("the three field tensors must have the same shape")
    if tuple(voxel_centers.shape) != tuple(slice_fields.shape[:3]) + (3,):
        raise RuntimeError("voxel_centers must be [Dx,Dy,Dz,3]")
    dev = slice_fields.device
    Dx, Dy, Dz = slice_fields.shape[:3]
    Ns, N1, N2 = slice_iso_values.numel(), lattice1_iso_values.numel(), lattice2_iso_values.numel()
    mask = torch.empty((Ns, N1, N2), dtype=torch.bool, device=dev)
    idx = torch.empty((Ns, N1, N2, 3), dtype=torch.int16, device=dev)
    pts = torch.empty((Ns, N1, N2, 3), dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws_bytes = L.inf3dp_fields_intersection_workspace_bytes(Dx, Dy, Dz, Ns, N1, N2)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.inf3dp_fields_intersection(*[_lib.ptr(t) for t in args], Dx, Dy, Dz, Ns, N1, N2, _lib.ptr(mask),
                                                _lib.ptr(idx), _lib.ptr(pts), _lib.ptr(ws), ws_bytes,
                                                _lib.stream_ptr(dev)))
        if sync:
            torch.cuda.current_stream(dev).synchronize()  # intersection.cu:281 cudaDeviceSynchronize
    return [mask, idx, pts]
