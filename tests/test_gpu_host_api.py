"""GPU parity tests of the HOST-FACING query API -- the calls a user of the reference makes and the one bench.py's `e2e`
number goes through: inf3dp.HostPipeline.run (pinned host buffers, three streams, double-buffered staging),
inf3dp.field_query_with_grads / field_querying (fused equivalents of utils.py:140-176, :217-266) and the sharding helpers
of inf3dp.dist at world size 1.  Every result is compared bit for bit with one device-resident launch of the same engine
(`net.query_rows` / `net.query`), and that launch with the fp64 oracle within the north-star tolerances."""
import numpy as np
import pytest
import torch

from conftest import layers_from_golden
from oracle import siren_oracle as so

pytestmark = pytest.mark.gpu

TOL_F = 1e-4 * np.sqrt(12.0)   # |df| <= 1e-4 of the bbox diagonal of [-1,1]^3
TOL_ANG = 0.1                  # degrees


def _net(golden, name="sdf", out_features=1):
    import inf3dp
    net = inf3dp.SingleBVPNet(type="sine", in_features=3, out_features=out_features)
    net.load_state_dict({k[len(name) + 1:]: torch.from_numpy(np.asarray(golden[k])) for k in golden.files
                         if k.startswith(name + "/net.")})
    return net.cuda()


def _points(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand((n, 3), generator=g) * 2 - 1).float()


@pytest.mark.parametrize("n,chunk", [(3 * 4096 + 777, 4096), (5 * 1000 + 1, 1000), (4096, 4096), (100, 4096),
                                     (7 * 12800 + 129, 12800)])
def test_host_pipeline_rows_equal_the_device_resident_query(golden, n, chunk):
    """>= 3 chunks so that both staging buffers are recycled (events ev_k / ev_out), N not a multiple of the chunk, a
    chunk that is not a multiple of the 128-point tile, N smaller than one chunk."""
    import inf3dp
    net = _net(golden)
    x = _points(n, seed=n)
    xh = x.pin_memory()
    oh = torch.full((n, 4), float("nan")).pin_memory()
    pipe = inf3dp.HostPipeline(net, chunk=chunk)
    for rep in range(2):    # the second run re-enters with the events of the first still recorded
        oh.fill_(float("nan"))
        pipe.run(xh, oh)
        rows = net.query_rows(x.cuda()).cpu()
        assert torch.equal(oh, rows), (n, chunk, rep)
    # and the rows are the north-star-grade answer
    layers = layers_from_golden(golden, "sdf")
    m = min(n, 2048)
    yo, Jo, _ = so.siren_jet(layers, x[:m].numpy().astype(np.float64), order=1)
    assert np.abs(oh[:m, 0].numpy() - yo[:, 0]).max() <= TOL_F
    assert so.angle_deg(oh[:m, 1:].numpy(), Jo[:, 0]).max() <= TOL_ANG


def test_host_pipeline_forward_mode_engine_and_unpinned_buffers(golden):
    """mode='tc_precise' takes the (y, J) -> packed-row copy path; pageable host tensors must work too (slower)."""
    import inf3dp
    net = _net(golden)
    n = 3 * 2048 + 5
    x = _points(n, seed=3)
    oh = torch.empty((n, 4))
    inf3dp.HostPipeline(net, chunk=2048, mode="tc_precise").run(x, oh)
    y, J, _ = net.query(x.cuda(), order=1, mode="tc_precise")
    assert torch.equal(oh[:, 0], y[:, 0].cpu()) and torch.equal(oh[:, 1:], J[:, 0].cpu())


@pytest.mark.parametrize("n,max_batch", [(10_000, 4096), (3 * 4096, 4096), (129, 1 << 22), (1, 7)])
def test_field_query_with_grads_and_field_querying_chunking(golden, n, max_batch):
    import inf3dp
    net = _net(golden)
    x = _points(n, seed=11).cuda()
    f, g = inf3dp.field_query_with_grads(net, x, max_batch=max_batch)
    y1, J1, _ = net.query(x, order=1)
    assert f.shape == (n,) and g.shape == (n, 3)
    assert torch.equal(f, y1[:, 0]) and torch.equal(g, J1[:, 0])
    v = inf3dp.field_querying(net, x, max_batch=max_batch)
    y0, _, _ = net.query(x, order=0)
    assert v.shape == (n,) and torch.equal(v, y0[:, 0])

    class Decoder(torch.nn.Module):          # the reference's wrapper pattern (exp_scripts/toolpath.py:21-30)
        def __init__(self, m):
            super().__init__()
            self.model = m

        def forward(self, coords):
            return self.model({"coords": coords})
    f2, g2 = inf3dp.field_query_with_grads(Decoder(net), x, max_batch=max_batch)
    assert torch.equal(f2, f) and torch.equal(g2, g)


def test_field_querying_multi_output_network(golden):
    import inf3dp
    quat = _net(golden, "quat", 4)
    x = _points(5000, seed=5).cuda()
    q = inf3dp.field_querying(quat, x, max_batch=2048)
    y, _, _ = quat.query(x, order=0)
    assert q.shape == (5000, 4) and torch.equal(q, y)
    yo, _, _ = so.siren_jet(layers_from_golden(golden, "quat"), x[:512].cpu().numpy().astype(np.float64), order=0)
    assert np.abs(q[:512].cpu().numpy() - yo).max() <= TOL_F


def test_sharding_helpers_at_world_size_one(golden):
    """inf3dp.dist on a single rank (no process group): the sharded entry points reduce to the plain ones, and the
    pipelined gather's staging rows are the rows of the query (gloo world-2 runs of the same code: tests/test_dist_gloo.py)."""
    import inf3dp
    from inf3dp import dist as idist
    from inf3dp.testing import torus_mesh
    net = _net(golden)
    n = 10_001
    x = _points(n, seed=2).cuda()
    rows = idist.sharded_value_grad(net, x)
    y, J, _ = net.query(x, order=1)
    assert torch.equal(rows[:, 0], y[:, 0]) and torch.equal(rows[:, 1:], J[:, 0])
    pipe = idist.PipelinedAllGather(n, 4, chunks=3, device=x.device)
    for c, (lo, hi) in enumerate(pipe.ranges):
        net.query_rows(x[lo:hi], out=pipe.slot(c))
        pipe.gather(c)
    pipe.finish()
    assert torch.equal(pipe.rows_of_rank(0), net.query_rows(x))
    V, F = torus_mesh(64, 32)
    f = (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0])).astype(np.float32)
    iso = np.linspace(f.min(), f.max(), 9).astype(np.float32)
    t = [torch.from_numpy(a).cuda() for a in (V, F, f, iso)]
    m, k = idist.sharded_isocontours(*t)
    m0, k0 = inf3dp.extract_isocontours(*t)
    assert torch.equal(m, m0) and torch.equal(k, k0)


def test_slice_waypoint_and_curvature_wrappers_at_world_size_one(golden):
    """inf3dp.dist.sharded_fields_intersection / sharded_quat_collision_sweep / sharded_curvature_sweep on a single rank are
    the plain device calls, bit for bit (their N > 1 exchange: tests/test_dist_gloo.py on gloo, bench.py --gpus N on NCCL)."""
    import copy
    import inf3dp
    from inf3dp import dist as idist
    from inf3dp.testing import corner_expanded_fields
    dev = torch.device("cuda:0")
    (sf, af, bf), centers = corner_expanded_fields(12)
    isos = (np.linspace(-0.9, 0.9, 17).astype(np.float32), np.linspace(-0.8, 0.8, 7).astype(np.float32),
            np.linspace(-0.8, 0.8, 5).astype(np.float32))
    t = [torch.from_numpy(a).to(dev) for a in (sf, af, bf, *isos, centers)]
    m0, i0, p0 = inf3dp.get_fields_intersection_inter_slice(*t)
    m1, i1, p1 = idist.sharded_fields_intersection(*t, sync=True)
    assert torch.equal(m0, m1) and torch.equal(i0, i1) and torch.equal(p0.nan_to_num(), p1.nan_to_num())
    assert int(m0.sum()) > 0

    net = _net(golden)
    x = _points(5000, seed=5).cuda()
    _, J, H = net.query(x, order=2)
    k0, d0 = inf3dp.principal_curvatures(J[:, 0], H[:, 0])
    k1, d1 = idist.sharded_curvature_sweep(net, x)
    k2, d2 = idist.sharded_curvature_sweep(net, x, presharded=True, n_total=5000)
    assert torch.equal(k0, k1) and torch.equal(d0, d1) and torch.equal(k0, k2) and torch.equal(d0, d2)
    with pytest.raises(RuntimeError, match="its shard of"):
        idist.sharded_curvature_sweep(net, x, presharded=True, n_total=4999)

    torch.manual_seed(2048)
    others = [inf3dp.SingleBVPNet(type="sine", in_features=3, out_features=o).to(dev) for o in (1, 1, 1, 1, 4)]
    lvl = inf3dp.MLPClassifier(num_classes=8).to(dev)
    sdf = copy.deepcopy(net)
    with torch.no_grad():
        sdf.net.net[4][0].bias -= 0.06
    way = x[:300].clamp(-0.5, 0.5)
    noz = (torch.rand((16, 3), generator=torch.Generator().manual_seed(1)) * 0.1).to(dev)
    lev = torch.arange(300, device=dev) % 8
    msl = torch.zeros((300, 8), device=dev)
    a = inf3dp.quat_collision_sweep(way, lev, msl, noz, others[4], sdf, others[1], lvl, 8, -0.9)
    b = idist.sharded_quat_collision_sweep(way, lev, msl, noz, others[4], sdf, others[1], lvl, 8, -0.9)
    c = idist.sharded_quat_collision_sweep(way, lev, msl, noz, others[4], sdf, others[1], lvl, 8, -0.9, presharded=True,
                                           n_total=300, waypoint_chunk=128)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[1], c[1])
    assert float((a[0] - c[0]).abs().max()) <= 1e-6
