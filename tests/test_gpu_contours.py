"""GPU parity tests (through the C ABI) for extract_isocontours / get_fields_intersection_inter_slice:
CUDA path vs the CPU oracle (bit-exact for counts, indices and -- in the oracle's face-ascending schedule --
the vertices themselves), vs the compiled reference extension when oracle/_ref is present, and vs the golden
files recorded from that extension on a B200 (tests/golden/contour_ref_*.npz, tests/golden/make_contour_golden.py).
A test that needs the real reference kernels FAILS when neither the extension nor its golden is available."""
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


def _reference_or_none():
    from oracle import build_ref
    return build_ref.load() if os.path.exists(build_ref.so_path()) else None


def _golden_or_none(name):
    p = os.path.join(GOLDEN, name)
    return np.load(p) if os.path.exists(p) else None


def _compare_with_reference_rows(m, n, s, ref_rows_of, rn):
    """Ours (dense rows m, loop counts n, segment counts s) vs ONE outcome of the reference kernels.  The reference's slot
    order is atomicAdd-dependent, so loops are compared as point sets (rotation / direction / which vertex the seed drops
    differ): identical segment counts S_k, identical loop counts on this clean fixture, every reference vertex within 1e-5
    of a vertex of ours except the one vertex per loop the reference's seed drops."""
    for k in range(len(n)):
        theirs_rows = ref_rows_of(k)
        assert len(theirs_rows) - rn[k] == s[k], k                       # identical segment count S_k
        if rn[k] == n[k] and len(theirs_rows):
            ours = np.concatenate(co.split_rows(m[k], n[k]) or [np.zeros((0, 3), np.float32)])
            theirs = theirs_rows[theirs_rows[:, 0] > -1e2]
            if len(ours) * len(theirs) < 4e7:
                d = np.abs(theirs[:, None, :] - ours[None, :, :]).max(-1).min(-1)
                assert (d > 1e-5).sum() <= rn[k], k
    assert (rn == n).mean() > 0.9


def test_vs_compiled_reference_extension_and_its_golden():
    """The real reference kernels on the same inputs: live (oracle/_ref, built by oracle/build_ref.py, travels with the
    gpurun snapshot) AND from the golden recorded from them on a B200 (tests/golden/contour_ref_isocontours.npz,
    tests/golden/make_contour_golden.py).  Whichever is present is used; if neither is, the test FAILS -- the only
    comparison against the real reference kernels must not disappear silently."""
    ref = _reference_or_none()
    gold = _golden_or_none("contour_ref_isocontours.npz")
    assert ref is not None or gold is not None, \
        "neither oracle/_ref (python oracle/build_ref.py) nor tests/golden/contour_ref_isocontours.npz is available"
    if gold is not None:
        V, f, iso = gold["V"], gold["fields"], gold["iso"]
        _, F = co.torus_mesh(int(gold["nu"]), int(gold["nv"]))
        m, n, s = _run(V, F, f, iso)
        offs = np.concatenate([[0], np.cumsum(gold["rows_used"])])
        assert np.array_equal(gold["seg_counts"], s)
        _compare_with_reference_rows(m, n, s, lambda k: gold["rows"][offs[k]:offs[k + 1]], gold["contour_nums"])
    if ref is not None:
        V, F = co.torus_mesh(160, 80)
        f = _field(V)
        iso = np.linspace(f.min(), f.max(), 40).astype(np.float32)
        tv, tf, tfl, ti = (torch.from_numpy(a).cuda() for a in (V, F, f, iso))
        rm, rn = ref.extract_isocontours(tv, tf, tfl, ti)
        rm, rn = rm.cpu().numpy(), rn.cpu().numpy()
        m, n, s = _run(V, F, f, iso)
        _compare_with_reference_rows(m, n, s, lambda k: rm[k, :co.rows_used(rm[k], rn[k])], rn)


def _grid_fields(D, seed=0):
    from inf3dp.testing import corner_expanded_fields
    return corner_expanded_fields(D, seed=seed, nan_frac=0.2)


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


def test_intersection_vs_compiled_reference_extension_and_its_golden():
    import inf3dp
    from inf3dp.testing import corner_expanded_fields
    ref = _reference_or_none()
    gold = _golden_or_none("contour_ref_intersection.npz")
    assert ref is not None or gold is not None, \
        "neither oracle/_ref (python oracle/build_ref.py) nor tests/golden/contour_ref_intersection.npz is available"
    (s, a, b), centers = corner_expanded_fields(24, poly=True)
    si = np.linspace(-0.9, 0.9, 40).astype(np.float32)
    ai = np.linspace(-0.8, 0.8, 12).astype(np.float32)
    bi = np.linspace(-0.8, 0.8, 12).astype(np.float32)
    t = [torch.from_numpy(x).cuda() for x in (s, a, b, si, ai, bi, centers)]
    m, i, p = [x.cpu().numpy() for x in inf3dp.get_fields_intersection_inter_slice(*t)]
    assert m.any()
    assert co.intersection_unmatched(s, a, b, si, ai, bi, centers, m, p, tol=1e-5) == 0
    outcomes = []
    if gold is not None:
        outcomes.append((gold["small_mask"], gold["small_idx"], gold["small_pts"]))
    if ref is not None:
        outcomes.append(tuple(x.cpu().numpy() for x in ref.get_fields_intersection_inter_slice(*t)))
    for rm, ri, rp in outcomes:
        assert np.array_equal(rm, m) and np.array_equal(ri, i)
        # the reference is last-writer-wins: its point must be the point of SOME voxel hitting the cell (oracle check),
        # ours is the lowest-voxel one; both are valid race outcomes
        assert co.intersection_unmatched(s, a, b, si, ai, bi, centers, rm, rp, tol=1e-5) == 0


def test_intersection_at_the_reference_size():
    """a6 at the size the reference itself uses (slice_toolpath.py:125): D = 255 voxels per side (256^3 samples), Ns = 500,
    N1 = N2 = 128 -- 16.6 M voxels, 8.2 M cells, every cell hit by ~1.5 voxels.  Mask and index triples bit-exact vs the C
    oracle (lowest-voxel winner rule), points <= 1e-6; vs the reference's own kernels (live and/or golden): mask and index
    triples bit-exact, and every point -- ours and theirs -- is the point of SOME voxel that hits its cell (candidate
    checker over all 8.2 M cells), because the reference's winner is a race."""
    import hashlib
    import inf3dp
    from inf3dp.testing import corner_expanded_fields
    ref = _reference_or_none()
    gold = _golden_or_none("contour_ref_intersection.npz")
    assert ref is not None or gold is not None, \
        "neither oracle/_ref (python oracle/build_ref.py) nor tests/golden/contour_ref_intersection.npz is available"
    (s, a, b), centers = corner_expanded_fields(255, nan_frac=0.0, poly=True)
    si = np.linspace(-0.9, 0.9, 500).astype(np.float32)
    ai = np.linspace(-0.8, 0.8, 128).astype(np.float32)
    bi = np.linspace(-0.8, 0.8, 128).astype(np.float32)
    t = [torch.from_numpy(x).cuda() for x in (s, a, b, si, ai, bi, centers)]
    mt, it, pt = inf3dp.get_fields_intersection_inter_slice(*t)
    m, i, p = mt.cpu().numpy(), it.cpu().numpy(), pt.cpu().numpy()
    mo, io, po = co.fields_intersection(s, a, b, si, ai, bi, centers, winner_rule=0)
    assert mo.sum() > 8_000_000
    assert np.array_equal(m, mo) and np.array_equal(i, io)
    assert np.array_equal(np.isnan(p), np.isnan(po)) and np.nanmax(np.abs(p - po)) <= 1e-6
    if gold is not None:
        assert hashlib.sha256(np.ascontiguousarray(m).tobytes()).hexdigest() == str(gold["big_mask_sha"])
        assert hashlib.sha256(np.ascontiguousarray(i).tobytes()).hexdigest() == str(gold["big_idx_sha"])
        # sampled points of the reference's run: the race may have picked another of the voxels that hit the cell (they
        # are neighbours: <= two voxel diagonals away); most cells have a single candidate and agree to rounding
        # (measured on the B200: 88 % of the sampled cells agree to 1e-5, the largest difference is 1.04 voxel diagonals)
        d = np.abs(p.reshape(-1, 3)[gold["big_cells"]] - gold["big_pts"]).max(1)
        assert d.max() <= 2 * 2.0 / 255 * np.sqrt(3.0) and (d <= 1e-5).mean() >= 0.5, (d.max(), (d <= 1e-5).mean())
    if ref is not None:
        rm, ri, rp = ref.get_fields_intersection_inter_slice(*t)
        assert torch.equal(rm, mt) and torch.equal(ri, it)
        assert co.intersection_unmatched(s, a, b, si, ai, bi, centers, m, rp.cpu().numpy(), tol=1e-5) == 0
    # grid mode (no corner expansion) gives the same answer at this size
    from inf3dp.testing import smooth_grids
    grids = [torch.from_numpy(g.astype(np.float32)).cuda() for g in smooth_grids(256, poly=True)]
    mg, ig, pg = inf3dp.get_fields_intersection_from_grids(*grids, *t[3:6])
    assert torch.equal(mg, mt) and torch.equal(ig, it)
    assert (pg - pt).abs().max().item() <= 1e-6


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


def test_many_isovalues_above_the_default_shared_memory_and_pool_overflow_retry():
    """K = 8192 isovalues: the face kernels' isovalue table (8 K bytes = 64 KB) exceeds the 48 KB default dynamic shared
    memory (the kernels must be opted in), and on this coarse mesh the segment total (~0.6 M) exceeds the
    default pool of nF + 65536 segments, so the first attempt reports INF3DP_EOVERFLOW and the retry is sized from the
    count pass (inf3dp_extract_isocontours_required_segments) -- not for the worst case nF * K."""
    import inf3dp
    V, F = co.torus_mesh(40, 20)
    f = _field(V)
    K = 8192
    iso = np.linspace(f.min(), f.max(), K).astype(np.float32)
    m, n, s = _run(V, F, f, iso)
    assert s.sum() > len(F) + 65536
    sel = np.arange(0, K, 257)
    mo, no, so_ = co.extract_isocontours(V, F, f, iso[sel])
    assert np.array_equal(s[sel], so_)
    for j, k in enumerate(sel):
        if n[k] == no[j]:
            assert np.array_equal(m[k], mo[j]), k
        else:
            assert n[k] < no[j], k
    assert (n[sel] == no).mean() >= 0.9
    # the compact entry point takes the same path
    rows, offs, counts, nums = inf3dp.extract_isocontours_compact(*[torch.from_numpy(a).cuda() for a in (V, F, f, iso)])
    assert np.array_equal(nums.cpu().numpy(), n) and np.array_equal(counts.cpu().numpy(), s + n)
    # beyond the documented limit: a clean error, no launch
    with pytest.raises(RuntimeError, match="12000"):
        _run(V, F, f, np.linspace(f.min(), f.max(), 12001).astype(np.float32))


def test_non_manifold_edge_three_faces_on_one_edge():
    """Three triangles share the edge (0,1) (a 'book' with three pages); spare non-crossing faces give the dense [K,F,3]
    contract room for S_k + C_k rows.  An isovalue crossing the shared edge puts THREE segment endpoints on it.  The
    reference chains by proximity: the walk's tail takes the lowest-index UNUSED segment with an endpoint within 1e-6
    (isocontours.cu:160-200).  Our (isovalue, mesh-edge) hash holds two endpoints per key (those of the two lowest faces)
    and links exactly that pair.
      iso 0.25  tails arrive on the shared edge from page 0 -> pages 0 + 1 joined, page 2 alone: identical (2 pieces)
      iso 0.70  the shared point is every seed's FIRST endpoint, which tail-only chaining never extends: identical (3)
      iso 0.52  page 1's tail arrives on the shared edge when page 0 (its hash partner) is already used; the reference's
                search skips it and joins page 2, ours ends the piece: 3 pieces instead of 2 -- DOCUMENTED DIVERGENCE
                (divergence 6; welded manifold meshes, e.g. every marching-cubes output, never
                have three faces on an edge).  Segment counts are identical in every case and the result is deterministic."""
    V = np.array([[0, 0, 0], [0, 0, 1], [1, 0, 0.5], [-0.5, 1, 0.5], [-0.5, -1, 0.5]] + [[5 + i, 5, 5] for i in range(9)], np.float32)
    F = np.array([[2, 0, 1], [3, 0, 1], [4, 0, 1]] + [[5 + 3 * i, 6 + 3 * i, 7 + 3 * i] for i in range(3)], np.int32)
    f = np.array([0.0, 1.0, 0.45, 0.55, 0.5] + [9.0] * 9, np.float32)
    iso = np.array([0.25, 0.7, 0.52], np.float32)
    mo, no, so_ = co.extract_isocontours(V, F, f, iso)
    assert list(no) == [2, 3, 2] and list(so_) == [3, 3, 3]
    first = None
    for rep in range(5):            # the hash insertion must not make the outcome timing dependent
        m, n, s = _run(V, F, f, iso)
        assert np.array_equal(s, so_)
        assert list(n) == [2, 3, 3], (rep, n)
        assert np.array_equal(m[:2], mo[:2]), rep
        # iso 0.52: three one-segment pieces, each = (seed's first endpoint, separator)
        assert co.rows_used(m[2], 3) == 6 and [len(p) for p in co.split_rows(m[2], 3)] == [1, 1, 1]
        first = m if first is None else first
        assert np.array_equal(m, first)
