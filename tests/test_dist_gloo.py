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
