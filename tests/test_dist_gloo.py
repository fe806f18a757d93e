"""CPU: the N>1 sharding / all-gather path on the gloo backend, world_size 2."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n):
    sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from inf3dp.dist import all_gather_rows, shard_range
    full = torch.arange(n * 4, dtype=torch.float32).view(n, 4)
    lo, hi, per = shard_range(n, rank, world)
    got = all_gather_rows(full[lo:hi].clone(), n)
    assert torch.equal(got, full), (rank, got.shape)
    cnt = all_gather_rows(torch.arange(lo, hi, dtype=torch.int32), n)
    assert torch.equal(cnt, torch.arange(n, dtype=torch.int32))
    # dtypes NCCL has no collective for travel as bytes: int16 index triples and bool masks of the intersection op
    idx16 = (torch.arange(n * 6, dtype=torch.int32).view(n, 2, 3) - 7).to(torch.int16)
    assert torch.equal(all_gather_rows(idx16[lo:hi].clone(), n), idx16)
    mask = (torch.arange(n * 5).view(n, 5) % 3 == 0)
    got_mask = all_gather_rows(mask[lo:hi].clone(), n)
    assert got_mask.dtype == torch.bool and torch.equal(got_mask, mask)
    dist.destroy_process_group()


def _pipe_worker(rank, world, port, n_local, chunks):
    sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from inf3dp.dist import PipelinedAllGather
    pipe = PipelinedAllGather(n_local, 4, chunks=chunks)
    mine = (torch.arange(n_local * 4, dtype=torch.float32).view(n_local, 4) + 1000.0 * rank)
    for c, (lo, hi) in enumerate(pipe.ranges):
        pipe.submit(c, mine[lo:hi, :1], mine[lo:hi, 1:])      # value column + gradient columns, as the bench does
    pipe.finish()
    for r in range(world):
        want = torch.arange(n_local * 4, dtype=torch.float32).view(n_local, 4) + 1000.0 * r
        assert torch.equal(pipe.rows_of_rank(r), want), (rank, r)
    # in-place producer: write the staging rows of a piece directly, then gather it
    for c, (lo, hi) in enumerate(pipe.ranges):
        pipe.slot(c).copy_(-mine[lo:hi])
        pipe.gather(c)
    pipe.finish()
    for r in range(world):
        want = -(torch.arange(n_local * 4, dtype=torch.float32).view(n_local, 4) + 1000.0 * r)
        assert torch.equal(pipe.rows_of_rank(r), want), (rank, r)
    dist.destroy_process_group()


def test_pipelined_all_gather_world2():
    for n_local, chunks in ((13, 4), (16, 8), (5, 8)):
        mp.spawn(_pipe_worker, args=(2, 29561 + n_local, n_local, chunks), nprocs=2, join=True)


def test_pipelined_all_gather_single_process():
    from inf3dp.dist import PipelinedAllGather
    pipe = PipelinedAllGather(10, 4, chunks=3)
    mine = torch.arange(40, dtype=torch.float32).view(10, 4)
    for c, (lo, hi) in enumerate(pipe.ranges):
        pipe.submit(c, mine[lo:hi, :1], mine[lo:hi, 1:])
    assert torch.equal(pipe.finish().shape and pipe.rows_of_rank(0), mine)


def test_shard_ranges_cover_exactly():
    from inf3dp.dist import shard_range
    for n in (0, 1, 7, 200, 16777216):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w)[:2] for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_all_gather_rows_world2():
    for n in (7, 8):
        mp.spawn(_worker, args=(2, 29533 + n, n), nprocs=2, join=True)


# --- the three other splits of SURVEY.md 8e (slice isovalues, waypoints, Hessian points): the sharding, index fix-up and exchange
# of inf3dp.dist.sharded_* on gloo.  The device calls are replaced by deterministic CPU stand-ins with the same signatures
# and output layouts (the real calls are compared with these wrappers at world size 1 in the GPU tests and at N > 1 by bench.py).

def _fake_intersection(sf, af, bf, s_iso, a_iso, b_iso, centers, sync=True):
    Ns, N1, N2 = s_iso.numel(), a_iso.numel(), b_iso.numel()
    key = (s_iso * 1000).round().long().view(Ns, 1, 1) + 3 * torch.arange(N1).view(1, N1, 1) + 5 * torch.arange(N2).view(1, 1, N2)
    mask = key % 3 == 0
    idx = torch.stack(torch.meshgrid(torch.arange(Ns), torch.arange(N1), torch.arange(N2), indexing="ij"), -1).to(torch.int16)
    idx = torch.where(mask.unsqueeze(-1), idx, torch.full_like(idx, -1))
    pts = torch.stack([s_iso.view(Ns, 1, 1).expand(Ns, N1, N2), a_iso.view(1, N1, 1).expand(Ns, N1, N2),
                       b_iso.view(1, 1, N2).expand(Ns, N1, N2)], -1)
    pts = torch.where(mask.unsqueeze(-1), pts, torch.full_like(pts, float("nan")))
    return [mask, idx, pts]


def _fake_collision_sweep(w, lev, msl, nozzle, quat, sdf, slc, level, num_classes, base_th, waypoint_chunk=1 << 18):
    d = (w.sum(1) * (lev.float() + 1.0) + msl[:, 0]) * float(nozzle.shape[0])
    return d.clamp_min(0.0), d > 0


class _FakeNet:
    def query(self, x, order=1, mode=None):
        assert order == 2
        J = torch.stack([x, 2 * x], 1)[:, :1]
        H = (x.unsqueeze(2) * x.unsqueeze(1)).unsqueeze(1)
        return x[:, :1].clone(), J, H


def _fake_curvatures(g, h):
    kap = torch.stack([g.sum(1), h.diagonal(dim1=1, dim2=2).sum(1)], 1)
    dirs = torch.stack([g, -g], 1)
    return kap, dirs


def _splits_worker(rank, world, port, Ns, n_way, n_pts):
    sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import inf3dp
    from inf3dp import contours, collision, diffops
    from inf3dp import dist as idist
    contours.get_fields_intersection_inter_slice = _fake_intersection
    collision.quat_collision_sweep = _fake_collision_sweep
    diffops.principal_curvatures = _fake_curvatures
    g = torch.Generator().manual_seed(7)
    # slice-isovalue split: global slice indices, -1 kept in empty cells, NaN points kept
    s_iso, a_iso, b_iso = torch.linspace(-0.9, 0.9, Ns), torch.linspace(-0.8, 0.8, 4), torch.linspace(-0.7, 0.7, 3)
    dummy = torch.zeros((2, 2, 2, 8))
    want = _fake_intersection(dummy, dummy, dummy, s_iso, a_iso, b_iso, None)
    got = idist.sharded_fields_intersection(dummy, dummy, dummy, s_iso, a_iso, b_iso, torch.zeros((2, 2, 2, 3)))
    assert got[0].dtype == torch.bool and got[1].dtype == torch.int16
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]), rank
    assert torch.equal(got[2].isnan(), want[2].isnan()) and torch.equal(got[2].nan_to_num(), want[2].nan_to_num())
    lo, hi, _ = idist.shard_range(Ns, rank, world)
    loc = idist.sharded_fields_intersection(dummy, dummy, dummy, s_iso, a_iso, b_iso, torch.zeros((2, 2, 2, 3)), gather=False)
    assert torch.equal(loc[1], want[1][lo:hi]) and torch.equal(loc[0], want[0][lo:hi])
    # waypoint split, replicated and presharded inputs
    way = torch.rand((n_way, 3), generator=g) - 0.5
    lev = torch.randint(0, 8, (n_way,), generator=g)
    msl = torch.rand((n_way, 8), generator=g) * 0.1
    noz = torch.zeros((5, 3))
    wd, wh = _fake_collision_sweep(way, lev, msl, noz, *(None,) * 4, 8, -0.9)
    d, h = idist.sharded_quat_collision_sweep(way, lev, msl, noz, None, None, None, None, 8, -0.9)
    assert h.dtype == torch.bool and torch.equal(d, wd) and torch.equal(h, wh), rank
    lo, hi, _ = idist.shard_range(n_way, rank, world)
    d, h = idist.sharded_quat_collision_sweep(way[lo:hi], lev[lo:hi], msl[lo:hi], noz, None, None, None, None, 8, -0.9,
                                              presharded=True, n_total=n_way, waypoint_chunk=3)
    assert torch.equal(d, wd) and torch.equal(h, wh), rank
    try:
        bad = hi - lo + 1
        idist.sharded_quat_collision_sweep(torch.zeros((bad, 3)), torch.zeros(bad, dtype=torch.long), torch.zeros((bad, 8)), noz,
                                           None, None, None, None, 8, -0.9, presharded=True, n_total=n_way)
        raise AssertionError("a shard of the wrong length was accepted")
    except RuntimeError as e:
        assert "its shard of" in str(e)
    # Hessian-point split
    x = torch.rand((n_pts, 3), generator=g) * 2 - 1
    _, J, H = _FakeNet().query(x, order=2)
    wk, wdirs = _fake_curvatures(J[:, 0], H[:, 0])
    k, dirs = idist.sharded_curvature_sweep(_FakeNet(), x)
    assert k.shape == (n_pts, 2) and dirs.shape == (n_pts, 2, 3) and torch.equal(k, wk) and torch.equal(dirs, wdirs), rank
    lo, hi, _ = idist.shard_range(n_pts, rank, world)
    k, dirs = idist.sharded_curvature_sweep(_FakeNet(), x[lo:hi], presharded=True, n_total=n_pts)
    assert torch.equal(k, wk) and torch.equal(dirs, wdirs), rank
    dist.destroy_process_group()


def test_slice_waypoint_and_hessian_point_splits_world2():
    # ragged shards, and more ranks than units (rank 1's shard is empty)
    for j, (Ns, n_way, n_pts) in enumerate(((7, 11, 9), (8, 16, 32), (1, 1, 1))):
        mp.spawn(_splits_worker, args=(2, 29611 + j, Ns, n_way, n_pts), nprocs=2, join=True)


def test_sharded_wrappers_single_process_pass_through():
    """Without an initialised process group the wrappers are the plain calls (world 1, nothing gathered)."""
    sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
    from inf3dp import contours, collision, diffops
    from inf3dp import dist as idist
    saved = (contours.get_fields_intersection_inter_slice, collision.quat_collision_sweep, diffops.principal_curvatures)
    contours.get_fields_intersection_inter_slice = _fake_intersection
    collision.quat_collision_sweep = _fake_collision_sweep
    diffops.principal_curvatures = _fake_curvatures
    try:
        s_iso, a_iso, b_iso = torch.linspace(-0.9, 0.9, 5), torch.linspace(-0.8, 0.8, 4), torch.linspace(-0.7, 0.7, 3)
        dummy = torch.zeros((2, 2, 2, 8))
        want = _fake_intersection(dummy, dummy, dummy, s_iso, a_iso, b_iso, None)
        got = idist.sharded_fields_intersection(dummy, dummy, dummy, s_iso, a_iso, b_iso, None)
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        x = torch.rand((6, 3))
        k, dirs = idist.sharded_curvature_sweep(_FakeNet(), x)
        assert k.shape == (6, 2) and dirs.shape == (6, 2, 3)
    finally:
        contours.get_fields_intersection_inter_slice, collision.quat_collision_sweep, diffops.principal_curvatures = saved


# --- isovalue split (dense and compact exchange) and point split of the plain query, on gloo with CPU stand-ins

def _fake_compact(vertices, faces, fields, iso_values):
    """Row streams like extract_isocontours_compact: isovalue v owns n(v) rows filled with v, streams in isovalue order, the row
    buffer larger than what is used (as the real one: capacity, not size)."""
    K = iso_values.numel()
    counts = ((iso_values * 10).round().long().abs() % 5 + 1).to(torch.int32) if K else torch.zeros(0, dtype=torch.int32)
    offs = torch.cumsum(counts.long(), 0) - counts.long()
    rows = torch.full((int(counts.sum()) + 7, 3), -7.0)
    for k in range(K):
        rows[offs[k]: offs[k] + counts[k]] = iso_values[k]
    nums = (counts // 2).to(torch.int32)
    return rows, offs, counts, nums


def _fake_dense(vertices, faces, fields, iso_values, sync=True, want_seg_counts=False):
    K, F = iso_values.numel(), faces.shape[0]
    merged = iso_values.view(K, 1, 1).expand(K, F, 3).contiguous() if K else torch.zeros((0, F, 3))
    return merged, (iso_values * 100).round().to(torch.int32)


class _FakeFieldNet:
    def query(self, x, order=1, mode=None):
        return x.sum(1, keepdim=True), (2 * x).unsqueeze(1), None


def _iso_worker(rank, world, port, K, n):
    sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from inf3dp import contours
    from inf3dp import dist as idist
    contours.extract_isocontours_compact = _fake_compact
    contours.extract_isocontours_raw = _fake_dense
    V, F = torch.zeros((4, 3)), torch.zeros((6, 3), dtype=torch.int32)
    iso = torch.linspace(-0.83, 0.91, K)
    rows, offs, counts, nums = idist.sharded_isocontours_compact(V, F, torch.zeros(4), iso)
    w_rows, w_offs, w_counts, w_nums = _fake_compact(V, F, None, iso)
    assert offs.dtype == torch.int64 and counts.dtype == torch.int32
    assert torch.equal(counts, w_counts) and torch.equal(nums, w_nums), rank
    assert rows.shape[0] == int(w_counts.sum())                      # only used rows travelled, no padding in between
    for k in range(K):                                               # every stream sits where its global offset says
        got = rows[offs[k]: offs[k] + counts[k]]
        assert got.shape[0] == int(w_counts[k]) and bool((got == iso[k]).all()), (rank, k)
    assert torch.equal(offs, w_offs)                                 # contiguous shards => the same packing as one rank's
    merged, dn = idist.sharded_isocontours(V, F, torch.zeros(4), iso)
    wm, wn = _fake_dense(V, F, None, iso)
    assert torch.equal(merged, wm) and torch.equal(dn, wn), rank
    x = torch.rand((n, 3), generator=torch.Generator().manual_seed(3))
    got = idist.sharded_value_grad(_FakeFieldNet(), x)
    assert torch.equal(got, torch.cat([x.sum(1, keepdim=True), 2 * x], 1)), rank
    dist.destroy_process_group()


def test_isovalue_split_dense_and_compact_and_point_split_world2():
    for j, (K, n) in enumerate(((7, 13), (8, 16), (1, 1))):
        mp.spawn(_iso_worker, args=(2, 29641 + j, K, n), nprocs=2, join=True)
