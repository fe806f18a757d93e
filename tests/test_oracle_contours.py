"""CPU: properties of the contour / intersection oracle (oracle/contour_oracle.c)."""
import numpy as np
import pytest

from oracle import contour_oracle as co


def _field(V):
    return (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)


def test_rows_equal_segments_plus_loops_and_loops_close():
    V, F = co.torus_mesh(96, 48)
    f = _field(V)
    iso = np.linspace(f.min(), f.max(), 24).astype(np.float32)
    merged, nums, segs = co.extract_isocontours(V, F, f, iso)
    assert merged.shape == (24, len(F), 3)
    assert nums[0] == 0 and nums[-1] == 0  # min / max isovalue touch a vertex only: no segment (isocontours.cu:83)
    for k in range(len(iso)):
        used = co.rows_used(merged[k], nums[k])
        assert used == segs[k] + nums[k]
        assert not merged[k, used:].any()  # zero padding
        for loop in co.split_rows(merged[k], nums[k]):
            if len(loop) > 8:
                # a closed loop: first point ~ far endpoint of the last segment (toolpath_utils.py:76)
                assert np.linalg.norm(loop[0] - loop[-1]) < 1e-4 or nums[k] > 1


def test_isovalue_on_vertex_yields_no_segment_there():
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5], [6, 5, 5], [5, 6, 5]], np.float32)
    F = np.array([[0, 1, 2], [3, 4, 5]], np.int32)  # second face far outside every isovalue
    f = np.array([0.0, 1.0, 2.0, 9.0, 9.0, 9.0], np.float32)
    merged, nums, segs = co.extract_isocontours(V, F, f, np.array([1.0, 0.5, 0.0, 2.0, 3.0], np.float32))
    assert list(segs) == [0, 1, 0, 0, 0]  # iso == f1 gives one crossing only -> dropped
    assert list(nums) == [0, 1, 0, 0, 0]
    assert np.allclose(merged[1, 0], [0.5, 0, 0]) and np.allclose(merged[1, 1], [-1e9] * 3)


def test_open_strip_splits_at_seed():
    # strip of 4 triangles crossed by one isoline: an open chain; the seed (lowest face) sits at one end here
    V = np.array([[0, 0, 0], [0, 1, 0], [1, 0, 0], [1, 1, 0], [2, 0, 0], [2, 1, 0], [0, 9, 0], [1, 9, 0], [0, 8, 0]], np.float32)
    F = np.array([[0, 2, 1], [1, 2, 3], [2, 4, 3], [3, 4, 5]] + [[6, 7, 8]] * 4, np.int32)  # 4 far faces = row room
    f = V[:, 1].copy()
    merged, nums, segs = co.extract_isocontours(V, F, f, np.array([0.5], np.float32))
    assert segs[0] == 4 and nums[0] >= 1
    assert co.rows_used(merged[0], nums[0]) == 4 + nums[0]


def test_empty_inputs():
    V, F = co.torus_mesh(8, 8)
    merged, nums, segs = co.extract_isocontours(V, F, _field(V), np.zeros(0, np.float32))
    assert merged.shape == (0, len(F), 3) and len(nums) == 0


def _grid_fields(D, seed=0):
    rng = np.random.default_rng(seed)
    N = D + 1
    g = np.stack(np.meshgrid(*[np.linspace(-1, 1, N)] * 3, indexing="ij"), -1)
    base = [g[..., 2] + 0.1 * np.sin(3 * g[..., 0]), g[..., 0] + 0.2 * g[..., 1] ** 2, g[..., 1] - 0.15 * g[..., 2] * g[..., 0]]
    offs = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]  # toolpath_utils.py:1002-1005
    out = []
    for b in base:
        c = np.stack([b[o[0]:o[0] + D, o[1]:o[1] + D, o[2]:o[2] + D] for o in offs], -1).astype(np.float32)
        out.append(c)
    vs = 2.0 / D
    idx = np.stack(np.meshgrid(*[np.arange(D)] * 3, indexing="ij"), -1).astype(np.float32)
    centers = ((idx + 0.5) * vs - 1.0).astype(np.float32)  # toolpath_utils.py:1027-1037
    # knock out a corner region with NaN, like the SDF mask does (toolpath_utils.py:1071-1084)
    nanmask = rng.random((D, D, D)) < 0.2
    for c in out:
        c[nanmask] = np.nan
    return out, centers


def test_intersection_winner_rules_and_candidate_check():
    (s, a, b), centers = _grid_fields(12)
    si = np.linspace(-0.9, 0.9, 17).astype(np.float32)
    ai = np.linspace(-0.8, 0.8, 7).astype(np.float32)
    bi = np.linspace(-0.8, 0.8, 5).astype(np.float32)
    m0, i0, p0 = co.fields_intersection(s, a, b, si, ai, bi, centers, winner_rule=0)
    m1, i1, p1 = co.fields_intersection(s, a, b, si, ai, bi, centers, winner_rule=1)
    assert m0.any() and np.array_equal(m0, m1) and np.array_equal(i0, i1)
    assert (i0[~m0] == -1).all() and (p0[~m0] == 0).all()
    ii = np.argwhere(m0)
    assert np.array_equal(i0[m0], ii.astype(np.int16))
    # points lie in the voxel grid; NaN is a legitimate reference outcome (0/0 on a flat edge, intersection.cu:147)
    assert np.nanmax(np.abs(p0[m0])) <= 1.0 + 1e-5
    assert not np.array_equal(np.nan_to_num(p0), np.nan_to_num(p1))  # some cell is hit by more than one voxel
    for m, p in ((m0, p0), (m1, p1)):
        assert co.intersection_unmatched(s, a, b, si, ai, bi, centers, m, p, tol=0.0) == 0
    bad = p0.copy()
    first_finite = next(tuple(c) for c in ii if np.isfinite(p0[tuple(c)]).all())
    bad[first_finite] += 0.01
    assert co.intersection_unmatched(s, a, b, si, ai, bi, centers, m0, bad, tol=1e-5) == 1
