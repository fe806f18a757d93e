"""Volume-side producers of the contour path, device resident.

  * grid_queries / grid_field / get_grid_fields   -- the reference's grid sweep (sdf_meshing.py:46-65 gen_voxle_queries +
    utils.py:140-176 field_querying, toolpath_utils.py:974-983 get_grid_fields) evaluated chunk by chunk on the device:
    the [N^3,3] coordinate tensor is never built, every chunk is generated (bit-identical to the reference, shear
    included) and consumed by one fused SIREN launch
  * marching_cubes / sdf_to_mesh / gen_iso_layers -- skimage.measure.marching_cubes's role at sdf_meshing.py:399 and
    utils.py:322 on the device
  * get_fields_intersection_from_grids             -- get_fields_intersection_inter_slice on the sample grids themselves:
    replaces extract_corner_fields (toolpath_utils.py:986-1019, an 8x expansion of each grid), get_voxel_centers
    (:1021-1033) and the extension call (ext.cpp:20-46) with one native call, bit-identical outputs

All of it goes through the C ABI in include/inf3dp.h; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .query import _net_of


def _require_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (inf3dp has no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def grid_queries(N, cube_size=1.0, start=0, count=None, device=None, out=None):
    """Rows [start, start+count) of the reference's gen_voxle_queries(N, cube_size) tensor, on the device."""
    total = int(N) ** 3
    count = total - start if count is None else int(count)
    dev = torch.device(device if device is not None else (out.device if out is not None else torch.cuda.current_device()))
    if dev.type != "cuda":
        raise RuntimeError("inf3dp.grid_queries needs a CUDA device")
    if out is None:
        out = torch.empty((count, 3), dtype=torch.float32, device=dev)
    else:
        _require_cuda_f32(out, "out")
        if out.numel() < 3 * count:
            raise RuntimeError("grid_queries: `out` is too small")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().inf3dp_grid_queries(int(N), float(cube_size), int(start), count, _lib.ptr(out),
                                                  _lib.stream_ptr(dev)))
    return out[:count] if out.dim() == 2 else out


@torch.no_grad()
def grid_field(decoder, N=256, cube_size=1.0, chunk=1 << 22, with_grads=False, device=None, mode=None):
    """Field (and optionally gradient) of a SIREN on the reference's N^3 query grid, device resident.

    -> fields [N,N,N] f32 (and grads [N,N,N,3] f32).  Equivalent to the reference's
    `field_querying(decoder, gen_voxle_queries(N)[0]).reshape(N,N,N)` (sdf_meshing.py:75-76, toolpath_utils.py:977-981).
    """
    net = _net_of(decoder)
    if net.out_features != 1:
        raise RuntimeError("grid_field: single-output field networks only")
    dev = torch.device(device if device is not None else next(net.parameters()).device)
    total = int(N) ** 3
    chunk = int(min(chunk, total))
    coords = torch.empty((chunk, 3), dtype=torch.float32, device=dev)
    fields = torch.empty(total, dtype=torch.float32, device=dev)
    grads = torch.empty((total, 3), dtype=torch.float32, device=dev) if with_grads else None
    for h in range(0, total, chunk):
        m = min(chunk, total - h)
        grid_queries(N, cube_size, h, m, out=coords)
        y, J, _ = net.query(coords[:m], order=1 if with_grads else 0, mode=mode)
        fields[h:h + m] = y[:, 0]
        if with_grads:
            grads[h:h + m] = J[:, 0, :]
    fields = fields.view(N, N, N)
    return (fields, grads.view(N, N, N, 3)) if with_grads else fields


def get_grid_fields(decoder_list, N=256):
    """toolpath_utils.py:974-983 -> list of [N,N,N] device tensors."""
    return [grid_field(d, N) for d in decoder_list]


def marching_cubes(volume, level=0.0, spacing=(1.0, 1.0, 1.0), gradient_direction="descent", return_normals=True):
    """skimage.measure.marching_cubes(volume, level=, spacing=) on the device.

    volume [Gx,Gy,Gz] f32 CUDA -> (verts [nv,3] f32, faces [nf,3] i32, normals [nv,3] f32, values [nv] f32), all on the
    device.  Raises ValueError when the level is outside the finite value range, as skimage does."""
    _require_cuda_f32(volume, "volume")
    if volume.dim() != 3:
        raise ValueError("marching_cubes: volume must be 3-dimensional")
    if gradient_direction not in ("descent", "ascent"):
        raise ValueError("gradient_direction must be 'descent' or 'ascent'")
    if len(spacing) != 3:
        raise ValueError("spacing must have three entries")
    dev = volume.device
    Gx, Gy, Gz = volume.shape
    L = _lib.lib()
    ws_bytes = L.inf3dp_marching_cubes_workspace_bytes(Gx, Gy, Gz)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    counts = (ctypes.c_int64 * 2)()
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        _lib.check(L.inf3dp_marching_cubes_count(_lib.ptr(volume), Gx, Gy, Gz, float(level), _lib.ptr(ws), ws_bytes,
                                                 ctypes.cast(counts, ctypes.c_void_p), st))
        nv, nf = int(counts[0]), int(counts[1])
        if nf == 0:
            raise ValueError("marching_cubes: no surface found at the given level")
        verts = torch.empty((nv, 3), dtype=torch.float32, device=dev)
        faces = torch.empty((nf, 3), dtype=torch.int32, device=dev)
        normals = torch.empty((nv, 3), dtype=torch.float32, device=dev) if return_normals else None
        values = torch.empty(nv, dtype=torch.float32, device=dev)
        _lib.check(L.inf3dp_marching_cubes_emit(_lib.ptr(volume), Gx, Gy, Gz, float(level), float(spacing[0]),
                                                float(spacing[1]), float(spacing[2]),
                                                1 if gradient_direction == "descent" else 0, _lib.ptr(verts),
                                                _lib.ptr(faces), _lib.ptr(normals), _lib.ptr(values), nv, nf,
                                                _lib.ptr(ws), ws_bytes, st))
    return verts, faces, normals, values


def sdf_to_mesh(decoder, N=256, level=0.0, offset=None, scale=None, chunk=1 << 22):
    """The compute of sdf_meshing.py:69-90 + :368-414 (create_mesh_volume -> convert_sdf_samples_to_ply) without the
    host: grid sweep, marching cubes, and the voxel -> world transform of :405-414.  -> (mesh_points [nv,3] f32,
    faces [nf,3] i32, sdf grid [N,N,N]) on the device."""
    sdf = grid_field(decoder, N, chunk=chunk)
    voxel_size = 2.0 / (N - 1)
    verts, faces, _, _ = marching_cubes(sdf, level=level, spacing=(voxel_size,) * 3, return_normals=False)
    mesh_points = verts + (-1.0)            # voxel_grid_origin = (-1,-1,-1), sdf_meshing.py:48,407-409
    if scale is not None:
        mesh_points = mesh_points / scale
    if offset is not None:
        mesh_points = mesh_points - torch.as_tensor(offset, dtype=torch.float32, device=mesh_points.device)
    return mesh_points, faces, sdf


def gen_iso_layers(fields, num_iso_layers, mask, N=128):
    """utils.py:305-331 on the device: NaN-masked field, evenly spaced isovalues, one mesh per isovalue.
    -> {iso_value: (verts, faces)} with device tensors; isovalues without a surface are skipped as in the reference."""
    _require_cuda_f32(fields, "fields")
    masked = torch.where(mask, fields, torch.full((), float("nan"), device=fields.device))
    valid = masked[~torch.isnan(masked)]
    iso_values = np.linspace(valid.min().item(), valid.max().item(), num_iso_layers)
    layers = {}
    for iso in iso_values:
        try:
            verts, faces, _, _ = marching_cubes(masked.contiguous(), level=float(iso), spacing=(2.0 / N,) * 3,
                                                return_normals=False)
        except ValueError:
            continue
        layers[iso] = (verts, faces)
    return layers


def get_fields_intersection_from_grids(slice_grid, lattice1_grid, lattice2_grid, slice_iso_values, lattice1_iso_values,
                                       lattice2_iso_values, sync=True):
    """get_fields_intersection_inter_slice (ext.cpp:20-46) fed by the [N,N,N] grids directly.

    Bit-identical to `get_fields_intersection_inter_slice(*map(extract_corner_fields, grids), isos...,
    get_voxel_centers(N))` (toolpath_utils.py:986-1033, :1120-1127) without the 8x expansion or the centre tensor.
    -> [inter_mask bool [Ns,N1,N2], isovalue_indices int16 [Ns,N1,N2,3], intersections f32 [Ns,N1,N2,3]]."""
    grids = [slice_grid, lattice1_grid, lattice2_grid]
    isos = [slice_iso_values, lattice1_iso_values, lattice2_iso_values]
    for t, n in zip(grids + isos, ["slice_grid", "lattice1_grid", "lattice2_grid", "slice_iso_values",
                                   "lattice1_iso_values", "lattice2_iso_values"]):
        _require_cuda_f32(t, n)
    if slice_grid.dim() != 3 or lattice1_grid.shape != slice_grid.shape or lattice2_grid.shape != slice_grid.shape:
        raise RuntimeError("the three grids must be [Gx,Gy,Gz] tensors of the same shape")
    dev = slice_grid.device
    Gx, Gy, Gz = slice_grid.shape
    Ns, N1, N2 = (t.numel() for t in isos)
    mask = torch.empty((Ns, N1, N2), dtype=torch.bool, device=dev)
    idx = torch.empty((Ns, N1, N2, 3), dtype=torch.int16, device=dev)
    pts = torch.empty((Ns, N1, N2, 3), dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws_bytes = L.inf3dp_fields_intersection_workspace_bytes(Gx - 1, Gy - 1, Gz - 1, Ns, N1, N2)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.inf3dp_fields_intersection_grid(*[_lib.ptr(t) for t in grids + isos], Gx, Gy, Gz, Ns, N1, N2,
                                                     _lib.ptr(mask), _lib.ptr(idx), _lib.ptr(pts), _lib.ptr(ws), ws_bytes,
                                                     _lib.stream_ptr(dev)))
        if sync:
            torch.cuda.current_stream(dev).synchronize()
    return [mask, idx, pts]
