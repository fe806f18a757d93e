"""GPU parity of the volume-side producers through the C ABI:

  * grid queries        bit-exact vs the oracle (itself pinned to the reference's gen_voxle_queries)
  * grid sweep          fused chunked sweep == query of the materialised grid
  * marching cubes      bit-exact vs oracle/volume_oracle.c (vertices, faces, normals, values, order included);
                        size-independent properties at 384^3 (closed, oriented, Euler characteristic, vertex count)
  * intersection on grids == the expanded-input op on the reference's corner expansion + centres (bit-exact), and the oracle
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, layers_from_golden
from oracle import contour_oracle as co
from oracle import volume_oracle as vo

pytestmark = pytest.mark.gpu


def test_grid_queries_bit_exact():
    import inf3dp
    g = np.load(os.path.join(GOLDEN, "volume_golden.npz"))
    assert np.array_equal(inf3dp.grid_queries(37, device="cuda").cpu().numpy(), g["queries37"])
    assert np.array_equal(inf3dp.grid_queries(21, cube_size=0.5, device="cuda").cpu().numpy(), g["queries21_half"])
    q300 = inf3dp.grid_queries(300, device="cuda").cpu().numpy()
    assert np.array_equal(q300[g["queries300_rows"]], g["queries300"])
    for start, count in [(0, 5000), (300 ** 3 - 777, 777), (16_777_000, 4096)]:
        assert np.array_equal(inf3dp.grid_queries(300, start=start, count=count, device="cuda").cpu().numpy(),
                              vo.grid_queries(300, start=start, count=count))
    # the bench grid: rows beyond 2^24 and the last rows of 512^3 against the oracle, whole tensor against the torch
    # formulation that test_dropin pins to the reference
    for start, count in [(2 ** 24 - 100, 4096), (512 ** 3 - 3000, 3000), (100_000_123, 2048)]:
        assert np.array_equal(inf3dp.grid_queries(512, start=start, count=count, device="cuda").cpu().numpy(),
                              vo.grid_queries(512, start=start, count=count))
    from inf3dp.query import gen_voxel_queries
    full, _, _ = gen_voxel_queries(256, device="cuda")
    assert torch.equal(inf3dp.grid_queries(256, device="cuda"), full)
    assert inf3dp.grid_queries(5, start=125, count=0, device="cuda").shape == (0, 3)
    with pytest.raises(RuntimeError):
        inf3dp.grid_queries(5, start=120, count=10, device="cuda")


def _sdf_net(golden):
    import inf3dp
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    net.load_state_dict({k[4:]: torch.from_numpy(np.asarray(golden[k])) for k in golden.files if k.startswith("sdf/net.")})
    return net.cuda()


def test_grid_field_equals_query_of_the_materialised_grid(golden):
    import inf3dp
    from inf3dp.query import gen_voxel_queries
    net = _sdf_net(golden)
    N = 45
    coords, _, _ = gen_voxel_queries(N, device="cuda")
    y, J, _ = net.query(coords, order=1)
    f, gr = inf3dp.grid_field(net, N, chunk=20_000, with_grads=True)    # ragged chunks on purpose
    assert f.shape == (N, N, N) and gr.shape == (N, N, N, 3)
    # same kernels on the same coordinates; only the engine picked for a 20 000-point chunk may differ from the one
    # picked for the whole grid (both parity grade: |df| <= 3.46e-4)
    assert (f.reshape(-1) - y[:, 0]).abs().max().item() <= 2e-5
    assert (gr.reshape(-1, 3) - J[:, 0]).abs().max().item() <= 2e-4
    f0 = inf3dp.grid_field(net, N)
    assert (f0 - f).abs().max().item() <= 3.46e-4
    assert len(inf3dp.get_grid_fields([net, net], N=9)) == 2


def _grid(N):
    ax = np.linspace(-1, 1, N, dtype=np.float32)
    return np.meshgrid(ax, ax, ax, indexing="ij")


def _mc_both(vol, level, spacing, direction, misalign=False):
    import inf3dp
    dvol = torch.from_numpy(np.ascontiguousarray(vol)).cuda()
    if misalign:     # a contiguous volume whose storage starts 4 bytes off a 16-byte boundary: the unvectorised passes
        store = torch.empty(dvol.numel() + 1, dtype=torch.float32, device="cuda")
        store[1:].copy_(dvol.reshape(-1))
        dvol = store[1:].view(dvol.shape)
        assert dvol.data_ptr() % 16 == 4 and dvol.is_contiguous()
    v, f, n, val = inf3dp.marching_cubes(dvol, level, spacing, direction)
    vo_, fo, no, valo = vo.marching_cubes(vol, level, spacing, direction)
    assert v.dtype == torch.float32 and f.dtype == torch.int32
    assert np.array_equal(v.cpu().numpy(), vo_), "vertices"
    assert np.array_equal(f.cpu().numpy(), fo), "faces"
    assert np.array_equal(n.cpu().numpy(), no), "normals"
    assert np.array_equal(val.cpu().numpy(), valo), "values"
    return vo_, fo


@pytest.mark.parametrize("direction", ["descent", "ascent"])
def test_marching_cubes_bit_exact_smooth_fields(direction):
    X, Y, Z = _grid(48)
    sph = (np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.6).astype(np.float32)
    v, f = _mc_both(sph, 0.0, (2 / 47,) * 3, direction)
    assert vo.euler_characteristic(len(v), f) == 2
    tor = ((np.sqrt(X ** 2 + Y ** 2) - 0.55) ** 2 + Z ** 2 - 0.25 ** 2).astype(np.float32)
    v, f = _mc_both(tor[:, 3:-2, 5:-6], 0.013, (1.0, 2.0, 0.5), direction)      # non-cubic, anisotropic, level != 0
    assert vo.euler_characteristic(len(v), f) == 0


def test_marching_cubes_bit_exact_all_cases_and_masks():
    rng = np.random.default_rng(0)
    vol = np.full((23, 26, 31), 5.0, np.float32)
    vol[1:-1, 1:-1, 1:-1] = rng.standard_normal((21, 24, 29)).astype(np.float32)
    _mc_both(vol, 0.0, (1.0, 1.0, 1.0), "descent")
    _mc_both(vol[1:-1, 1:-1, 1:-1], 0.1, (1.0, 1.0, 1.0), "ascent")          # surface runs into the grid border
    X, Y, Z = _grid(40)
    sph = (np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.6).astype(np.float32)
    m = sph.copy()
    m[X > 0.2] = np.nan                                                     # utils.py:308 NaN mask
    m[(Y < -0.3) & (Z > 0.1)] = np.inf
    v, f = _mc_both(m, 0.0, (0.05,) * 3, "descent")
    assert np.isfinite(v).all() and len(np.unique(f)) == len(v)
    noise_nan = vol.copy()
    noise_nan[rng.random(vol.shape) < 0.05] = np.nan                        # scattered holes: the slow validity path
    _mc_both(noise_nan, 0.0, (1.0, 1.0, 1.0), "descent")
    _mc_both(sph[19:21, 19:21, :].copy(), 0.0, (1.0, 1.0, 1.0), "descent")      # a single column of cubes


def test_marching_cubes_degenerate_inputs():
    import inf3dp
    X, Y, Z = _grid(16)
    sph = torch.from_numpy((np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.6).astype(np.float32)).cuda()
    with pytest.raises(ValueError):
        inf3dp.marching_cubes(sph, 5.0)                     # skimage: "Surface level must be within volume data range"
    with pytest.raises(ValueError):
        inf3dp.marching_cubes(sph[:1].contiguous(), 0.0)    # no cubes
    with pytest.raises(RuntimeError):
        inf3dp.marching_cubes(sph.cpu(), 0.0)               # no CPU fallback
    with pytest.raises(RuntimeError):
        inf3dp.marching_cubes(sph.double(), 0.0)
    v, f, n, val = inf3dp.marching_cubes(sph, 0.0, return_normals=False)
    assert n is None and len(v) > 0


def test_marching_cubes_large_properties():
    """384^3 (57 M samples, 27 k tiles): vertex count == sign-changing edges, closed oriented surface, genus 0."""
    import inf3dp
    N = 384
    ax = torch.linspace(-1, 1, N, device="cuda")
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    vol = (torch.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.7 + 0.05 * torch.sin(9 * X) * torch.cos(7 * Y)).contiguous()
    del X, Y, Z
    v, f, n, val = inf3dp.marching_cubes(vol, 0.0, (2 / (N - 1),) * 3)
    inside = vol < 0
    crossing = sum(int((inside.narrow(a, 0, N - 1) != inside.narrow(a, 1, N - 1)).sum()) for a in range(3))
    assert len(v) == crossing
    f64 = f.long()
    assert int(f64.min()) == 0 and int(f64.max()) == len(v) - 1
    d = torch.cat([f64[:, [0, 1]], f64[:, [1, 2]], f64[:, [2, 0]]], 0)
    key = d[:, 0] * len(v) + d[:, 1]
    rkey = d[:, 1] * len(v) + d[:, 0]
    uk, ck = torch.unique(key, return_counts=True)
    assert int(ck.max()) == 1                                       # no directed edge twice: consistently oriented
    assert torch.equal(uk, torch.unique(rkey))                       # every directed edge has its reverse: closed
    E = len(uk) // 2
    assert len(v) - E + len(f) == 2
    # vertices lie on the level set of the analytic field to interpolation accuracy; normals agree with the winding
    p = v - 1.0
    val_at = torch.sqrt((p ** 2).sum(1)) - 0.7 + 0.05 * torch.sin(9 * p[:, 0]) * torch.cos(7 * p[:, 1])
    assert val_at.abs().max().item() < 5e-4
    tri = v[f64]
    fn = torch.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0], dim=1)
    vn = n[f64].mean(1)
    assert ((fn * vn).sum(1) > 0).float().mean().item() > 0.999


def test_sdf_to_mesh_and_iso_layers(golden):
    import inf3dp
    net = _sdf_net(golden)
    with torch.no_grad():
        net.net.net[4][0].bias -= 0.06       # the random-init SDF is positive everywhere: shift so a level set exists
    pts, faces, sdf = inf3dp.sdf_to_mesh(net, N=96, level=0.0)
    assert sdf.shape == (96, 96, 96) and pts.shape[1] == 3 and faces.shape[1] == 3
    assert pts.min().item() >= -1.0 - 1e-6 and pts.max().item() <= 1.0 + 1e-6
    y, _, _ = net.query(pts.contiguous(), order=0)
    assert y.abs().max().item() < 5e-2       # on the zero level up to the interpolation error of an omega=30 SIREN at N=96
    vo_, fo, _, _ = vo.marching_cubes(sdf.cpu().numpy(), 0.0, (2 / 95,) * 3)
    assert np.array_equal(faces.cpu().numpy(), fo)
    layers = inf3dp.gen_iso_layers(sdf, 5, sdf < 0.05, N=96)
    assert 1 <= len(layers) <= 5
    for iso, (v, f) in layers.items():
        assert v.is_cuda and f.shape[1] == 3 and torch.isfinite(v).all()


def _three_grids(N, seed=0, nan_frac=0.02):
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(*[np.linspace(-1, 1, N)] * 3, indexing="ij"), -1)
    grids = [g[..., 2] + 0.1 * np.sin(3 * g[..., 0]), g[..., 0] + 0.2 * g[..., 1] ** 2, g[..., 1] - 0.15 * g[..., 2] * g[..., 0]]
    grids = [np.ascontiguousarray(x.astype(np.float32)) for x in grids]
    hole = rng.random((N, N, N)) < nan_frac
    for x in grids:
        x[hole] = np.nan
    return grids


@pytest.mark.parametrize("N,Ns,N1,N2", [(13, 17, 7, 5), (32, 60, 16, 16), (9, 1, 1, 1), (2, 3, 2, 2)])
def test_intersection_on_grids_equals_the_expanded_path(N, Ns, N1, N2):
    import inf3dp
    grids = _three_grids(N)
    rng = np.random.default_rng(N)
    si = rng.permutation(np.linspace(-0.9, 0.9, Ns)).astype(np.float32)
    ai = np.linspace(-0.8, 0.8, N1).astype(np.float32)
    bi = np.linspace(-0.8, 0.8, N2).astype(np.float32)
    exp = [vo.corner_fields(x) for x in grids]            # pinned to the reference's extract_corner_fields
    centers = vo.voxel_centers(N)                         # pinned to the reference's get_voxel_centers
    tg = [torch.from_numpy(x).cuda() for x in grids + [si, ai, bi]]
    m, i, p = inf3dp.get_fields_intersection_from_grids(*tg)
    te = [torch.from_numpy(x).cuda() for x in exp + [si, ai, bi, centers]]
    me, ie, pe = inf3dp.get_fields_intersection_inter_slice(*te)
    assert torch.equal(m, me) and torch.equal(i, ie)
    assert torch.equal(torch.isnan(p), torch.isnan(pe))
    assert torch.equal(torch.nan_to_num(p), torch.nan_to_num(pe))        # bit-identical points
    mo, io, po = co.fields_intersection(*exp, si, ai, bi, centers, winner_rule=0)
    assert np.array_equal(m.cpu().numpy(), mo) and np.array_equal(i.cpu().numpy(), io)
    assert np.nanmax(np.abs(p.cpu().numpy() - po), initial=0.0) <= 1e-6


def test_marching_cubes_random_volumes_bit_exact():
    """Random shapes (all three count paths: Gz % 32 == 0 bit volume, Gz % 4 == 0 vectorised, scalar), levels, NaN / inf
    holes, both orientations: every array bit-identical to the oracle."""
    rng = np.random.default_rng(7)
    for trial in range(40):
        shape = [int(v) for v in rng.integers(2, 40, 3)]
        if trial % 2 == 0:
            shape[2] = 4 * max(1, shape[2] // 4)
        if trial % 4 == 1:
            shape[2] = 32 * int(rng.integers(1, 4))          # the bit-volume count path (Gz % 32 == 0)
        vol = rng.standard_normal(shape).astype(np.float32)
        if trial % 3 == 0:
            vol[rng.random(shape) < 0.05] = np.nan
        if trial % 5 == 0:
            vol[rng.random(shape) < 0.02] = np.inf
        level = float(rng.uniform(-0.5, 0.5))
        direction = "descent" if trial % 2 else "ascent"
        spacing = tuple(float(s) for s in rng.uniform(0.5, 2.0, 3))
        try:
            _mc_both(vol, level, spacing, direction, misalign=(trial % 8 in (1, 2)))
        except ValueError:
            assert len(vo.marching_cubes(vol, level, spacing, direction)[1]) == 0    # no surface: both agree
