"""GPU tests of the query kernel's peer-store epilogue (inf3dp_siren_value_grad_rows_peers): on ONE GPU the "peers" are
several local buffers -- every one of them must receive exactly the rows of the plain query.  The multi-GPU form
(symmetric memory, NVLink multicast) runs in bench.py --gpus N, which checks the gathered slabs against the local ones."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _net(golden):
    import inf3dp
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    net.load_state_dict({k[4:]: torch.from_numpy(np.asarray(golden[k])) for k in golden.files if k.startswith("sdf/net.")})
    return net.cuda()


@pytest.mark.parametrize("n,peers", [(5000, 1), (4097, 3), (128, 8), (70001, 8)])
def test_rows_land_in_every_peer_buffer(golden, n, peers):
    net = _net(golden)
    g = torch.Generator().manual_seed(n)
    x = (torch.rand((n, 3), generator=g) * 2 - 1).cuda()
    want = net.query_rows(x)
    bufs = [torch.full((n + 7, 4), float("nan"), device="cuda") for _ in range(peers)]
    # each destination starts 3 rows into its buffer: the pointer is the address of row 0 of THIS launch
    net.query_rows_peers(x, [b.data_ptr() + 3 * 16 for b in bufs])
    torch.cuda.synchronize()
    for b in bufs:
        assert torch.equal(b[3:3 + n], want)
        assert torch.isnan(b[:3]).all() and torch.isnan(b[3 + n:]).all()     # nothing outside the destination rows


def test_peer_argument_validation(golden):
    net = _net(golden)
    x = torch.zeros((10, 3), device="cuda")
    with pytest.raises(RuntimeError, match="peer"):
        net.query_rows_peers(x, [])
    buf = torch.zeros((10, 4), device="cuda")
    with pytest.raises(RuntimeError, match="peer"):
        net.query_rows_peers(x, [buf.data_ptr()] * 9)
    with pytest.raises(RuntimeError, match="aligned"):
        net.query_rows_peers(x, [buf.data_ptr() + 4])


def test_fused_row_gather_and_sharded_host_pipeline_single_rank_group(golden):
    """inf3dp.dist.FusedRowGather / ShardedHostPipeline on a ONE-rank NCCL group: symmetric-memory allocation, rendezvous,
    the (multicast or peer-store) epilogue and the barriers all run for real; the gathered slab must equal the plain query.
    (The multi-rank form is exercised and checked by `bench.py --gpus N`, which aborts on any mismatch.)"""
    import os
    import torch.distributed as dist
    from inf3dp import dist as idist
    net = _net(golden)
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", torch.cuda.current_device()))
        created = True
    try:
        n = 3 * 4096 + 123
        g = torch.Generator().manual_seed(5)
        x = (torch.rand((n, 3), generator=g) * 2 - 1)
        want = net.query_rows(x.cuda())
        for prefer in ("multicast", "p2p"):
            try:
                fg = idist.FusedRowGather(n, torch.device("cuda", torch.cuda.current_device()), prefer=prefer)
            except Exception as e:     # symmetric memory is a property of the box (driver / fabric), not of this code
                pytest.skip(f"symmetric memory unavailable on this box: {type(e).__name__}: {e}")
            fg.begin()
            fg.query(net, x.cuda())
            fg.finish()
            torch.cuda.synchronize()
            assert fg.mode in ("multicast", "p2p")
            assert torch.equal(fg.rows_of_rank(0), want), (prefer, fg.mode)
            # in two launches with a row offset (how ShardedHostPipeline feeds it)
            fg.gathered.zero_()
            fg.begin()
            fg.query(net, x[:5000].cuda(), row_offset=0)
            fg.query(net, x[5000:].cuda(), row_offset=5000)
            fg.finish()
            assert torch.equal(fg.local_rows(), want), prefer
            del fg
        hp = idist.ShardedHostPipeline(net, n, chunk=4096, device=torch.device("cuda", torch.cuda.current_device()))
        xh, oh = x.pin_memory(), torch.full((n, 4), float("nan")).pin_memory()
        for _ in range(2):
            hp.run(xh, oh, source_rank=0)
            assert torch.equal(oh, want.cpu())
    finally:
        if created:
            dist.destroy_process_group()
