"""CPU: the volume oracle (oracle/volume_oracle.c) against the goldens recorded from the reference's own grid helpers
(tests/golden/make_golden_volume.py), and the properties that define a correct marching-cubes output.

Marching cubes itself is PARITY UNPINNED against skimage (absent here and in /root/reference, see volume_oracle.c);
what is checked is the part every variant shares (one vertex per sign-changing edge, on the level set of the trilinear
edge interpolant) and the topology: closed, consistently oriented, right Euler characteristic."""
import hashlib
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import volume_oracle as vo

G = np.load(os.path.join(ROOT, "tests", "golden", "volume_golden.npz"))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_grid_queries_match_the_reference_bit_for_bit():
    assert np.array_equal(vo.grid_queries(37), G["queries37"])
    assert np.array_equal(vo.grid_queries(21, cube_size=0.5), G["queries21_half"])
    assert _sha(vo.grid_queries(128)) == str(G["queries128_sha256"])
    q300 = vo.grid_queries(300)          # indices above 2^24: float32 rounding of the true division matters
    assert np.array_equal(q300[G["queries300_rows"]], G["queries300"])
    assert _sha(q300) == str(G["queries300_sha256"])
    # chunked generation equals the corresponding rows
    assert np.array_equal(vo.grid_queries(300, start=26_999_000, count=1000), q300[26_999_000:])


def test_corner_fields_and_centers_match_the_reference():
    assert np.array_equal(vo.corner_fields(G["corner_grid"]), G["corner_fields"])
    assert np.array_equal(vo.voxel_centers(9), G["centers9"])
    c = vo.voxel_centers(256)
    assert np.array_equal(c[np.arange(255), np.arange(255), np.arange(255)], G["centers256_diag"])
    assert _sha(c) == str(G["centers256_sha256"])


def _grid(N):
    ax = np.linspace(-1, 1, N, dtype=np.float32)
    return np.meshgrid(ax, ax, ax, indexing="ij")


def _crossing_edges(vol, level):
    """Number of grid edges whose finite ends lie on different sides of the level (inside = value < level)."""
    n = 0
    inside = vol < level
    fin = np.isfinite(vol)
    for a in range(3):
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[a] = slice(0, -1)
        hi[a] = slice(1, None)
        n += int(((inside[tuple(lo)] != inside[tuple(hi)]) & fin[tuple(lo)] & fin[tuple(hi)]).sum())
    return n


@pytest.mark.parametrize("direction", ["descent", "ascent"])
def test_marching_cubes_sphere(direction):
    N = 40
    X, Y, Z = _grid(N)
    sdf = (np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.6).astype(np.float32)
    h = 2.0 / (N - 1)
    v, f, n, val = vo.marching_cubes(sdf, 0.0, (h,) * 3, direction)
    assert len(v) == _crossing_edges(sdf, 0.0)                      # one welded vertex per sign-changing edge
    edges, open_edges, dup_directed = vo.mesh_edge_stats(f)
    assert open_edges == 0 and dup_directed == 0                    # closed, consistently oriented 2-manifold
    assert vo.euler_characteristic(len(v), f) == 2
    r = np.linalg.norm(v - 1.0, axis=1)                             # verts are index*spacing: the cube is [0,2]^3
    assert np.abs(r - 0.6).max() < 2e-3                             # linear interpolation error of a curved field
    tri = v[f]
    fn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    outward = ((fn * (tri.mean(1) - 1.0)).sum(1) > 0)
    # the SDF grows outwards: 'ascent' faces point outwards, 'descent' (skimage's default) inwards
    assert outward.all() if direction == "ascent" else (~outward).all()
    cosn = (n * (v - 1.0)).sum(1) / r
    assert (cosn > 0.999).all() if direction == "ascent" else (cosn < -0.999).all()
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-6)
    assert (val >= 0).all() and (val < 2 * h).all()                 # the larger end of a crossing edge


def test_marching_cubes_torus_and_anisotropic_spacing():
    X, Y, Z = _grid(44)
    tor = ((np.sqrt(X ** 2 + Y ** 2) - 0.55) ** 2 + Z ** 2 - 0.25 ** 2).astype(np.float32)
    v, f, _, _ = vo.marching_cubes(tor[:, :, 4:-4], 0.0, (1.0, 2.0, 0.5))
    assert vo.mesh_edge_stats(f)[1:] == (0, 0)
    assert vo.euler_characteristic(len(v), f) == 0
    v1, f1, _, _ = vo.marching_cubes(tor[:, :, 4:-4], 0.0, (1.0, 1.0, 1.0))
    assert np.array_equal(f, f1) and np.allclose(v, v1 * np.array([1.0, 2.0, 0.5], dtype=np.float32), rtol=1e-6)


def test_marching_cubes_ambiguous_noise_is_closed():
    """White noise exercises all 256 cases and every ambiguous face: the surface must stay closed (every mesh edge in an
    even number of triangles; a fan diagonal lying in a shared cube face can be used by both cubes -> 4, never odd)."""
    rng = np.random.default_rng(0)
    vol = np.full((22, 22, 22), 5.0, np.float32)
    vol[1:-1, 1:-1, 1:-1] = rng.standard_normal((20, 20, 20)).astype(np.float32)
    seen = set()
    for x in range(21):
        for y in range(21):
            for z in range(21):
                c = vol[x:x + 2, y:y + 2, z:z + 2]
                seen.add(int(sum((1 << (i + 2 * j + 4 * k)) for i in range(2) for j in range(2) for k in range(2) if c[i, j, k] < 0)))
    assert len(seen) >= 250
    v, f, _, _ = vo.marching_cubes(vol, 0.0)
    assert len(v) == _crossing_edges(vol, 0.0)
    d = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0).astype(np.int64)
    u = np.sort(d, 1)
    _, cnt = np.unique(u[:, 0] * len(v) + u[:, 1], return_counts=True)
    assert set(np.unique(cnt)) <= {2, 4}
    # orientation is consistent: every directed edge is matched by its reverse the same number of times
    key = d[:, 0] * len(v) + d[:, 1]
    rkey = d[:, 1] * len(v) + d[:, 0]
    ka, ca = np.unique(key, return_counts=True)
    kb, cb = np.unique(rkey, return_counts=True)
    assert np.array_equal(ka, kb) and np.array_equal(ca, cb)


def test_marching_cubes_nan_mask_and_empty():
    X, Y, Z = _grid(32)
    sdf = (np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.6).astype(np.float32)
    m = sdf.copy()
    m[X > 0.2] = np.nan                                            # utils.py:308 masks fields with NaN
    v, f, n, val = vo.marching_cubes(m, 0.0)
    assert np.isfinite(v).all() and np.isfinite(n).all() and len(v) > 0
    assert len(np.unique(f)) == len(v)                             # no vertex is left unreferenced
    assert (v[:, 0] <= (np.searchsorted(X[:, 0, 0], 0.2) - 1) + 1e-6).all()   # nothing beyond the last finite slab
    _, open_edges, _ = vo.mesh_edge_stats(f)
    assert open_edges > 0                                          # the cut is an open boundary
    v0, f0, _, _ = vo.marching_cubes(sdf, 5.0)                     # level outside the value range
    assert len(v0) == 0 and len(f0) == 0
    v1, f1, _, _ = vo.marching_cubes(sdf[:1], 0.0)                 # a slab one sample thick has no cubes
    assert len(v1) == 0 and len(f1) == 0


def test_volume_api_has_no_cpu_fallback():
    """The product entry points reject host tensors (as CHECK_CUDA does in the reference, contour_cuda/includes/utils.h:6-8)
    and never route through the oracle."""
    import torch
    import inf3dp
    vol = torch.zeros((4, 4, 4))
    with pytest.raises(RuntimeError):
        inf3dp.marching_cubes(vol, 0.0)
    with pytest.raises(RuntimeError):
        inf3dp.get_fields_intersection_from_grids(vol, vol, vol, torch.zeros(2), torch.zeros(2), torch.zeros(2))
    with pytest.raises(RuntimeError):
        inf3dp.grid_queries(8, device="cpu")
    src = open(os.path.join(ROOT, "inf-3dp_b200", "inf3dp", "volume.py")).read()
    assert "oracle" not in src.replace("no CPU fallback", "")


def test_c_abi_rejects_bad_volume_arguments():
    from inf3dp import _lib
    L = _lib.lib()
    assert L.inf3dp_marching_cubes_workspace_bytes(0, 4, 4) == 0
    assert L.inf3dp_marching_cubes_workspace_bytes(64, 64, 64) >= 64 ** 3 * 6
    assert L.inf3dp_grid_queries(1, 1.0, 0, 0, None, None) == _lib.EINVAL              # N < 2
    assert L.inf3dp_grid_queries(8, 1.0, 500, 100, None, None) == _lib.EINVAL          # range outside the grid
    assert L.inf3dp_marching_cubes_count(None, 4, 4, 4, 0.0, None, 0, None, None) == _lib.EINVAL
    assert L.inf3dp_siren_set_sm_limit(-1) == _lib.EINVAL and L.inf3dp_siren_set_sm_limit(0) == _lib.OK


def test_marching_cubes_random_volumes_property():
    """Random small volumes (smooth + noise, random shapes, levels, NaN holes): one vertex per sign-changing edge between
    finite samples that touches a valid cube, every referenced vertex exists, every mesh edge away from holes and the grid
    border is shared by an even number of triangles, and 'ascent' is 'descent' with every triangle flipped."""
    rng = np.random.default_rng(42)
    for trial in range(40):
        shape = tuple(int(v) for v in rng.integers(2, 9, 3))
        vol = rng.standard_normal(shape).astype(np.float32)
        if trial % 3 == 0:
            vol[rng.random(shape) < 0.1] = np.nan
        level = float(rng.uniform(-0.5, 0.5))
        v, f, n, val = vo.marching_cubes(vol, level, (1.0, 1.0, 1.0), "descent")
        va, fa, na, vala = vo.marching_cubes(vol, level, (1.0, 1.0, 1.0), "ascent")
        assert np.array_equal(v, va) and np.array_equal(val, vala) and np.array_equal(n, -na)
        assert np.array_equal(f, fa[:, [0, 2, 1]])
        if len(f):
            assert f.min() >= 0 and f.max() < len(v) and len(np.unique(f)) == len(v)
        assert np.isfinite(v).all() and np.isfinite(n).all()
        assert len(v) <= _crossing_edges(vol, level)
        if not np.isnan(vol).any():
            # closed except on the grid border: pad with a far-outside shell and the surface must close completely
            pad = np.full(tuple(s + 2 for s in shape), 100.0, np.float32)
            pad[1:-1, 1:-1, 1:-1] = vol
            vp, fp, _, _ = vo.marching_cubes(pad, level)
            assert len(vp) == _crossing_edges(pad, level)
            d = np.concatenate([fp[:, [0, 1]], fp[:, [1, 2]], fp[:, [2, 0]]], 0).astype(np.int64)
            u = np.sort(d, 1)
            _, cnt = np.unique(u[:, 0] * max(len(vp), 1) + u[:, 1], return_counts=True)
            assert (cnt % 2 == 0).all()
