"""GPU parity tests (through the C ABI) for extract_isocontours / get_fields_intersection_inter_slice:
CUDA path vs the CPU oracle (bit-exact for counts, indices and -- in the oracle's face-ascending schedule --
the vertices themselves), vs the compiled reference extension when oracle/_ref is present, and vs the golden
files recorded from that extension on a B200 (tests/golden/contour_ref_*.npz)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import contour_oracle as co

pytestmark = pytest.mark.gpu


def _field(V):
    return (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)


def assert_equivalent(m, n, s, mo, no, so_):
    """Ours vs the oracle.  Segment counts always identical.  Where the loop counts agree the row streams are
    bit-identical.  They can differ only one way: the reference's proximity search (isocontours.cu:160-200, any
    unused endpoint within 1e-6) can hop over a tiny segment next to a mesh vertex and strand it as an extra piece
    (SURVEY.md appendix 10); the edge-keyed chaining keeps it in its loop, so ours has FEWER pieces and the same
    vertices (<= 1e-5)."""
    assert np.array_equal(s, so_)
    for k in range(len(n)):
        if n[k] == no[k]:
            assert np.array_equal(m[k], mo[k]), k
            continue
        assert n[k] < no[k], k
        mine = np.concatenate(co.split_rows(m[k], n[k]))
        theirs = np.concatenate(co.split_rows(mo[k], no[k]))
        d = np.abs(theirs[:, None, :] - mine[None, :, :]).max(-1)
        assert (d.min(1) <= 1e-5).all() and ((d.min(0) > 1e-5).sum() <= no[k]), k
    assert (n == no).mean() >= 0.9


def _run(V, F, f, iso):
    import inf3dp
    m, n, s = inf3dp.extract_isocontours_raw(torch.from_numpy(V).cuda(), torch.from_numpy(F).cuda(),
                                             torch.from_numpy(f).cuda(), torch.from_numpy(iso).cuda(), want_seg_counts=True)
    return m.cpu().numpy(), n.cpu().numpy(), s.numpy()


@pytest.mark.parametrize("nu,nv,K", [(64, 32, 20), (200, 100, 64), (37, 19, 7)])
def test_bit_exact_vs_oracle_on_torus(nu, nv, K):
    V, F = co.torus_mesh(nu, nv)
    f = _field(V)
    iso = np.linspace(f.min(), f.max(), K).astype(np.float32)
    mo, no, so_ = co.extract_isocontours(V, F, f, iso)
    m, n, s = _run(V, F, f, iso)
    # identical segment counts; identical vertices, order, separators and zero padding (bit-exact) wherever the
    # reference's proximity search does not strand a segment
    assert_equivalent(m, n, s, mo, no, so_)


def test_unsorted_and_duplicate_isovalues_and_vertex_hits():
    V, F = co.torus_mesh(48, 24)
    f = _field(V)
    rng = np.random.default_rng(1)
    iso = rng.uniform(f.min() - 0.1, f.max() + 0.1, 33).astype(np.float32)
    iso[5] = iso[17]              # duplicate isovalue
    iso[9] = f[123]               # exactly on a vertex value: that vertex's edges give one crossing, no segment
    mo, no, so_ = co.extract_isocontours(V, F, f, iso)
    m, n, s = _run(V, F, f, iso)
    assert_equivalent(m, n, s, mo, no, so_)


def test_open_mesh_and_tiny_and_empty():
    V, F = co.torus_mesh(40, 20)
    f = _field(V)
    keep = np.ones(len(F), bool)
    keep[100:400] = False           # tear a hole: open chains, several pieces per isovalue
    F2 = np.ascontiguousarray(F[keep])
    iso = np.linspace(f.min(), f.max(), 15).astype(np.float32)
    mo, no, so_ = co.extract_isocontours(V, F2, f, iso)
    m, n, s = _run(V, F2, f, iso)
    assert_equivalent(m, n, s, mo, no, so_)
    # K = 0
    m0, n0, _ = _run(V, F2, f, np.zeros(0, np.float32))
    assert m0.shape == (0, len(F2), 3) and n0.shape == (0,)
    # single triangle + spare faces
    Vt = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5], [6, 5, 5], [5, 6, 5]], np.float32)
    Ft = np.array([[0, 1, 2], [3, 4, 5]], np.int32)
    ft = np.array([0.0, 1.0, 2.0, 9.0, 9.0, 9.0], np.float32)
    it = np.array([1.0, 0.5, 0.0, 2.0, 3.0], np.float32)
    mo, no, so_ = co.extract_isocontours(Vt, Ft, ft, it)
    m, n, s = _run(Vt, Ft, ft, it)
    assert np.array_equal(s, so_) and np.array_equal(n, no) and np.array_equal(m, mo)


@pytest.mark.parametrize("nu,nv,K,drop,seed", [(120, 60, 40, 0.03, 0), (160, 80, 24, 0.002, 1), (60, 30, 12, 0.4, 2),
                                               (300, 150, 16, 0.01, 3)])
def test_randomly_torn_meshes_open_chains(nu, nv, K, drop, seed):
    """Faces removed at random: open chains of every length, several chains and many pieces per isovalue (the
    interval replay of the parallel walk; with 40 % of the faces gone also its capacity fallback to the serial walk)."""
    V, F = co.torus_mesh(nu, nv)
    f = _field(V)
    rng = np.random.default_rng(seed)
    F2 = np.ascontiguousarray(F[rng.random(len(F)) >= drop])
    iso = np.linspace(f.min(), f.max(), K).astype(np.float32)
    mo, no, so_ = co.extract_isocontours(V, F2, f, iso)
    m, n, s = _run(V, F2, f, iso)
    assert_equivalent(m, n, s, mo, no, so_)


@pytest.mark.parametrize("torn", [False, True])
def test_compact_rows_equal_the_used_prefix_of_the_dense_contract(torn):
    """extract_isocontours_compact / contour_loops: the compact row streams are bit-identical to the
    used rows of the dense [K,F,3] result, and the host-side loops equal the reference's separator split."""
    import inf3dp
    V, F = co.torus_mesh(96, 48)
    f = _field(V)
    if torn:
        rng = np.random.default_rng(5)
        F = np.ascontiguousarray(F[rng.random(len(F)) >= 0.02])
    iso = np.linspace(f.min(), f.max(), 31).astype(np.float32)
    t = [torch.from_numpy(a).cuda() for a in (V, F, f, iso)]
    m, n = inf3dp.extract_isocontours(*t)
    rows, offs, counts, nums = inf3dp.extract_isocontours_compact(*t)
    assert torch.equal(nums, n)
    m, n = m.cpu().numpy(), n.cpu().numpy()
    rows, offs, counts = rows.cpu().numpy(), offs.cpu().numpy(), counts.cpu().numpy()
    for k in range(len(iso)):
        used = co.rows_used(m[k], n[k])
        assert counts[k] == used
        assert np.array_equal(rows[offs[k]:offs[k] + used], m[k, :used])
    loops = inf3dp.contour_loops(*t)
    for k in range(len(iso)):
        ref = co.split_rows(m[k], n[k])
        assert len(loops[k]) == len(ref) and all(np.array_equal(a, b) for a, b in zip(loops[k], ref))


@pytest.mark.parametrize("drop", [0.0, 0.0004])
def test_multi_million_face_mesh_big_walk_tier(drop):
    """F = 4 M faces (SURVEY.md 8d cfg 3, upper size): 5 - 10 k segments per isovalue, i.e. the second shared-memory tier of
    the parallel walk (closed loops, and with faces removed open chains whose state ids need 15 bits).

    At this resolution a few segments per isovalue are shorter than the reference's 1e-6 merge radius; its proximity search
    (isocontours.cu:160-200) then strands them as extra 1-segment pieces or passes them through the other endpoint
    (DESIGN.md divergence 1), so the comparison is the north-star one: identical segment counts, never more pieces, every
    vertex within 1e-5 -- plus, where the piece counts agree, the SAME row order, and for our output on its own the
    chaining invariant (consecutive rows of a piece are the two ends of one segment; closed pieces close)."""
    from scipy.spatial import cKDTree
    V, F = co.torus_mesh(2000, 1000)
    f = _field(V)
    if drop > 0:
        rng = np.random.default_rng(17)
        F = np.ascontiguousarray(F[rng.random(len(F)) >= drop])
    lo, hi = float(f.min()), float(f.max())
    iso = np.linspace(lo + 0.08 * (hi - lo), hi - 0.08 * (hi - lo), 10).astype(np.float32)
    mo, no, so_ = co.extract_isocontours(V, F, f, iso)
    m, n, s = _run(V, F, f, iso)
    assert s.max() > 5120 and s.max() <= 10752, s          # the tier under test
    assert np.array_equal(s, so_)
    edge = 2 * np.pi * 0.9 / 1000 * 1.5                      # longest mesh edge (diagonal of a quad), generous
    for k in range(len(iso)):
        assert n[k] <= no[k], k
        used, used_o = co.rows_used(m[k], n[k]), co.rows_used(mo[k], no[k])
        assert used == s[k] + n[k] and not m[k, used:].any(), k
        mine, theirs = co.split_rows(m[k], n[k]), co.split_rows(mo[k], no[k])
        pm, pt = np.concatenate(mine), np.concatenate(theirs)
        assert (cKDTree(pm).query(pt)[0] <= 1e-5).all(), k                       # every reference vertex is ours too
        assert (cKDTree(pt).query(pm)[0] > 1e-5).sum() <= no[k], k               # ours: at most the dropped seed vertices extra
        if n[k] == no[k]:
            d = np.abs(m[k, :used] - mo[k, :used_o]).max(1)
            assert d.max() <= 1e-5 and (d > 0).sum() <= 8, k                     # same order, same vertices
        for loop in mine:
            if len(loop) > 1:
                assert np.linalg.norm(np.diff(loop, axis=0), axis=1).max() <= edge, k   # chained through shared segments
            if drop == 0 and len(loop) > 50:
                assert np.linalg.norm(loop[0] - loop[-1]) <= edge, k                   # closed mesh: loops close


def test_full_size_properties():
    """F = 0.5 M faces, K = 200: properties that hold at any size (no O(S^2) oracle here): rows used = S_k + C_k,
    zero padding, every long loop closes, determinism."""
    V, F = co.torus_mesh(1000, 250)
    f = _field(V)
    iso = np.linspace(f.min(), f.max(), 200).astype(np.float32)
    m, n, s = _run(V, F, f, iso)
    m2, n2, s2 = _run(V, F, f, iso)
    assert np.array_equal(m, m2) and np.array_equal(n, n2) and np.array_equal(s, s2)
    assert s.sum() > 100000
    for k in range(0, 200, 7):
        used = co.rows_used(m[k], n[k])
        assert used == s[k] + n[k]
        assert not m[k, used:].any()
        for loop in co.split_rows(m[k], n[k]):
            if len(loop) > 50:
                assert np.linalg.norm(loop[0] - loop[-1]) < 1e-4


def _canon_loops(merged_k, num):
    return co.split_rows(merged_k, num)


def test_vs_compiled_reference_extension():
    """The real reference kernels (oracle/_ref, built by oracle/build_ref.py) on the same inputs.  The reference's
    slot order is atomicAdd-dependent, so loops are compared as point sets (rotation / direction / which vertex
    the seed drops differ): identical segment counts, identical loop counts on this clean fixture, every reference
    vertex within 1e-5 of a vertex of ours."""
    from oracle import build_ref
    if not os.path.exists(build_ref.so_path()):
        pytest.skip("oracle/_ref not built")
    ref = build_ref.load()
    V, F = co.torus_mesh(160, 80)
    f = _field(V)
    iso = np.linspace(f.min(), f.max(), 40).astype(np.float32)
    tv, tf, tfl, ti = (torch.from_numpy(a).cuda() for a in (V, F, f, iso))
    rm, rn = ref.extract_isocontours(tv, tf, tfl, ti)
    rm, rn = rm.cpu().numpy(), rn.cpu().numpy()
    m, n, s = _run(V, F, f, iso)
    for k in range(len(iso)):
        ref_rows = co.rows_used(rm[k], rn[k])
        assert ref_rows - rn[k] == s[k]                      # identical segment count S_k
        if rn[k] == n[k]:
            ours = np.concatenate(_canon_loops(m[k], n[k]) or [np.zeros((0, 3), np.float32)])
            theirs = np.concatenate(_canon_loops(rm[k], rn[k]) or [np.zeros((0, 3), np.float32)])
            if len(theirs):
                d = np.abs(theirs[:, None, :] - ours[None, :, :]).max(-1).min(-1) if len(ours) * len(theirs) < 4e7 else None
                if d is not None:
                    # each loop drops one vertex (the seed's p2) and the dropped one depends on the seed
                    assert (d > 1e-5).sum() <= rn[k]
    assert (rn == n).mean() > 0.9


def _grid_fields(D, seed=0):
    rng = np.random.default_rng(seed)
    N = D + 1
    g = np.stack(np.meshgrid(*[np.linspace(-1, 1, N)] * 3, indexing="ij"), -1)
    base = [g[..., 2] + 0.1 * np.sin(3 * g[..., 0]), g[..., 0] + 0.2 * g[..., 1] ** 2, g[..., 1] - 0.15 * g[..., 2] * g[..., 0]]
    offs = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
    out = [np.ascontiguousarray(np.stack([b[o[0]:o[0] + D, o[1]:o[1] + D, o[2]:o[2] + D] for o in offs], -1).astype(np.float32)) for b in base]
    idx = np.stack(np.meshgrid(*[np.arange(D)] * 3, indexing="ij"), -1).astype(np.float32)
    centers = ((idx + 0.5) * (2.0 / D) - 1.0).astype(np.float32)
    nanmask = rng.random((D, D, D)) < 0.2
    for c in out:
        c[nanmask] = np.nan
    return out, centers


@pytest.mark.parametrize("D,Ns,N1,N2", [(12, 17, 7, 5), (31, 60, 16, 16), (8, 1, 1, 1)])
def test_intersection_vs_oracle(D, Ns, N1, N2):
    import inf3dp
    (s, a, b), centers = _grid_fields(D)
    rng = np.random.default_rng(D)
    si = rng.permutation(np.linspace(-0.9, 0.9, Ns)).astype(np.float32)  # unsorted on purpose
    ai = np.linspace(-0.8, 0.8, N1).astype(np.float32)
    bi = np.linspace(-0.8, 0.8, N2).astype(np.float32)
    mo, io, po = co.fields_intersection(s, a, b, si, ai, bi, centers, winner_rule=0)
    t = [torch.from_numpy(x).cuda() for x in (s, a, b, si, ai, bi, centers)]
    m, i, p = inf3dp.get_fields_intersection_inter_slice(*t)
    assert m.dtype == torch.bool and i.dtype == torch.int16
    m, i, p = m.cpu().numpy(), i.cpu().numpy(), p.cpu().numpy()
    assert np.array_equal(m, mo)            # bit-exact mask
    assert np.array_equal(i, io)            # bit-exact index triples (-1 fill included)
    assert np.array_equal(np.isnan(p), np.isnan(po))
    assert np.nanmax(np.abs(p - po), initial=0.0) <= 1e-6   # north_star vertex tolerance is 1e-5
    assert co.intersection_unmatched(s, a, b, si, ai, bi, centers, m, p, tol=1e-6) == 0


def test_intersection_vs_compiled_reference_extension():
    from oracle import build_ref
    import inf3dp
    if not os.path.exists(build_ref.so_path()):
        pytest.skip("oracle/_ref not built")
    ref = build_ref.load()
    (s, a, b), centers = _grid_fields(24)
    si = np.linspace(-0.9, 0.9, 40).astype(np.float32)
    ai = np.linspace(-0.8, 0.8, 12).astype(np.float32)
    bi = np.linspace(-0.8, 0.8, 12).astype(np.float32)
    t = [torch.from_numpy(x).cuda() for x in (s, a, b, si, ai, bi, centers)]
    rm, ri, rp = [x.cpu().numpy() for x in ref.get_fields_intersection_inter_slice(*t)]
    m, i, p = [x.cpu().numpy() for x in inf3dp.get_fields_intersection_inter_slice(*t)]
    assert np.array_equal(rm, m) and np.array_equal(ri, i)
    # the reference is last-writer-wins: its point must be the point of SOME voxel hitting the cell (oracle check),
    # ours is the lowest-voxel one; both are valid race outcomes
    assert co.intersection_unmatched(s, a, b, si, ai, bi, centers, rm, rp, tol=1e-5) == 0
    assert co.intersection_unmatched(s, a, b, si, ai, bi, centers, m, p, tol=1e-5) == 0


def test_intersection_nonuniform_isovalue_sets():
    """Isovalue sets that defeat the interpolation guess of the window search (clustered, duplicated, far outliers, one or
    two entries): results stay those of the oracle."""
    import inf3dp
    (s, a, b), centers = _grid_fields(14, seed=3)
    rng = np.random.default_rng(99)
    si = np.concatenate([rng.uniform(-0.9, 0.9, 20) ** 3, [0.25, 0.25, 0.25, -50.0, 80.0]]).astype(np.float32)
    rng.shuffle(si)
    ai = np.concatenate([np.full(3, 0.1), rng.normal(0, 0.02, 9), [1e6]]).astype(np.float32)
    for bi in (np.array([0.05], np.float32), np.array([-0.3, 0.3], np.float32), np.sort(rng.uniform(-0.8, 0.8, 7)).astype(np.float32)):
        mo, io, po = co.fields_intersection(s, a, b, si, ai, bi, centers, winner_rule=0)
        t = [torch.from_numpy(x).cuda() for x in (s, a, b, si, ai, bi, centers)]
        m, i, p = [x.cpu().numpy() for x in inf3dp.get_fields_intersection_inter_slice(*t)]
        assert np.array_equal(m, mo) and np.array_equal(i, io)
        assert np.array_equal(np.isnan(p), np.isnan(po))
        assert np.nanmax(np.abs(p - po), initial=0.0) <= 1e-6
        assert m.any()
