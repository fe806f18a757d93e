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
