"""GPU parity tests (through the C ABI) for the fused SIREN jet: CUDA path vs the numpy oracle (fp64) and the
golden vectors recorded from the reference.  Tolerances are the north_star ones, written out here:
    |df| <= 1e-4 * bbox diagonal (3.46e-4 for [-1,1]^3), gradient angle <= 0.1 degrees.
"""
import numpy as np
import pytest
import torch

from conftest import layers_from_golden
from oracle import siren_oracle as so

pytestmark = pytest.mark.gpu

TOL_F = 1e-4 * np.sqrt(12.0)
TOL_ANGLE = 0.1


def _net(golden, name="sdf", out=1):
    import inf3dp
    net = inf3dp.SingleBVPNet(type="sine", in_features=3, out_features=out)
    sd = {k[len(name) + 1:]: torch.from_numpy(np.asarray(golden[k])) for k in golden.files if k.startswith(name + "/net.")}
    net.load_state_dict(sd)
    return net.cuda()


ENGINES = ["simt", "tc_precise"]
ENGINES_GRAD1 = ENGINES + ["tc_reverse"]   # value+gradient of single-output networks: also the reverse-mode engine


@pytest.mark.parametrize("engine", ENGINES_GRAD1)
def test_value_gradient_vs_reference_golden(golden, engine):
    net = _net(golden)
    x = torch.from_numpy(golden["x"]).cuda()
    y, J, _ = net.query(x, order=1, mode=engine)
    y, J = y.cpu().numpy(), J.cpu().numpy()
    assert np.abs(y - golden["sdf/y"]).max() <= TOL_F
    assert so.angle_deg(J[:, 0], golden["sdf/grad"]).max() <= TOL_ANGLE
    # and against the fp64 oracle (the reference's own fp32 error is ~1e-7 / 0.007 deg)
    yo, Jo, _ = so.siren_jet(layers_from_golden(golden, "sdf"), golden["x"].astype(np.float64), order=1)
    assert np.abs(y - yo).max() <= TOL_F
    assert so.angle_deg(J[:, 0], Jo[:, 0]).max() <= TOL_ANGLE
    if engine == "simt":  # the FP32 engine is as accurate as the reference itself
        assert np.abs(y - yo).max() < 5e-6
        assert so.angle_deg(J[:, 0], Jo[:, 0]).max() < 0.02


@pytest.mark.parametrize("engine", ENGINES)
def test_value_only_and_quaternion_out4(golden, engine):
    net = _net(golden, "quat", 4)
    x = torch.from_numpy(golden["x"]).cuda()
    y0, J0, _ = net.query(x, order=0, mode=engine)
    assert J0 is None and np.abs(y0.cpu().numpy() - golden["quat/y"]).max() <= TOL_F
    y1, J1, _ = net.query(x, order=1, mode=engine)
    assert np.abs(y1.cpu().numpy() - golden["quat/y"]).max() <= TOL_F
    Jr = golden["quat/jac"]
    for c in range(4):
        assert so.angle_deg(J1.cpu().numpy()[:, c], Jr[:, c]).max() <= TOL_ANGLE


@pytest.mark.parametrize("engine", ["simt", "tc_precise", "tc_reverse", "auto"])
def test_hessian_and_curvature_vs_reference(golden, engine):
    """hessian3d / min_max_curvs_and_vectors of the reference (diff_operators.py:27-44, :167-220) vs the order-2 engines:
    FP32 CUDA cores, the tensor-core second-order forward jet (10 jet rows per point) and the forward-over-reverse sweep
    (4 rows per point, what AUTO runs)."""
    import inf3dp
    net = _net(golden)
    x = torch.from_numpy(golden["sdf/hess_x"]).cuda()
    y, J, H = net.query(x, order=2, mode=engine)
    Hn = H.cpu().numpy()
    ref = golden["sdf/hess3d"]
    rel = np.linalg.norm((Hn - ref).reshape(len(Hn), -1), axis=1) / np.linalg.norm(ref.reshape(len(ref), -1), axis=1)
    assert rel.max() <= 1e-3  # proposed tolerance; fp32-vs-fp64 floor of the reference: 1.5e-5
    assert np.array_equal(Hn, np.swapaxes(Hn, -1, -2))
    kap, dirs = inf3dp.principal_curvatures(J[:64, 0], H[:64, 0])
    kr, dr = golden["sdf/curv_kappa"], golden["sdf/curv_dirs"]
    kap, dirs = kap.cpu().numpy(), dirs.cpu().numpy()
    assert np.abs(kap - kr).max() <= 2e-2 * max(1.0, np.abs(kr).max())
    sep = np.abs(kr[:, 1] - kr[:, 0]) > 1e-2 * np.abs(kr).max()
    assert (np.abs((dirs * dr).sum(-1))[sep] > np.cos(np.radians(0.5))).all()


def test_reference_call_pattern_autograd(golden):
    """forward({'coords'}) -> dict; diff_operators.gradient / hessian3d / hessian on it (modules.py:143-160,
    diff_operators.py:6-44,128-132), for [N,3] and [1,N,3]."""
    import inf3dp
    net = _net(golden)
    x = torch.from_numpy(golden["x"][:256]).cuda()
    out = net({"coords": x})
    assert out["model_in"].requires_grad and out["model_in"].is_leaf and out["model_out"].shape == (256, 1)
    g = inf3dp.gradient(out["model_out"], out["model_in"])
    assert so.angle_deg(g.detach().cpu().numpy(), golden["sdf/grad"][:256]).max() <= TOL_ANGLE
    # the stock autograd formulation of the reference (no fused shortcut): grad of grad
    gy = torch.ones_like(out["model_out"][..., 0])
    dydx = torch.autograd.grad(out["model_out"][..., 0], out["model_in"], gy, create_graph=True)[0]
    row0 = torch.autograd.grad(dydx[..., 0], out["model_in"], gy, create_graph=True)[0]
    xs = torch.from_numpy(golden["sdf/hess_x"]).cuda()
    o2 = net({"coords": xs})
    h = inf3dp.hessian3d(o2["model_out"], o2["model_in"]).cpu().numpy()
    ref = golden["sdf/hess3d"]
    assert np.abs(h - ref).max() <= 1e-3 * np.abs(ref).max()
    dydx2 = torch.autograd.grad(o2["model_out"][..., 0], o2["model_in"], torch.ones_like(o2["model_out"][..., 0]), create_graph=True)[0]
    r0 = torch.autograd.grad(dydx2[..., 0], o2["model_in"], torch.ones_like(dydx2[..., 0]), create_graph=True)[0]
    assert np.abs(r0.detach().cpu().numpy() - ref[:, 0, 0, :]).max() <= 1e-3 * np.abs(ref).max()
    assert row0.shape == (256, 3)
    ob = net({"coords": xs[None]})
    assert ob["model_out"].shape == (1, len(xs), 1)
    hb, status = inf3dp.hessian(ob["model_out"], ob["model_in"])
    assert status == 0 and np.abs(hb.cpu().numpy()[0] - ref).max() <= 1e-3 * np.abs(ref).max()
    gb = inf3dp.gradient(ob["model_out"], ob["model_in"])
    assert gb.shape == (1, len(xs), 3)
    # custom grad_outputs (diff_operators.py:129-131)
    go = torch.full_like(out["model_out"], 2.0)
    g2 = inf3dp.gradient(out["model_out"], out["model_in"], go)
    assert torch.allclose(g2, 2 * g, rtol=1e-6, atol=1e-8)


def test_level_mlp_relu(golden):
    """exp_scripts/train_levels.py:35-51 logits: the tensor-core classifier engine (default) and the FP32 SIMT engine
    against the logits recorded from the reference; ragged sizes; C = 3 (the reference's default) and C = 8."""
    import inf3dp
    lvl = inf3dp.MLPClassifier(num_classes=8)
    lvl.load_state_dict({k[len("level/"):]: torch.from_numpy(np.asarray(golden[k])) for k in golden.files if k.startswith("level/network")})
    lvl.cuda()
    xg = torch.from_numpy(golden["level/x"]).cuda()
    lo = lvl(xg).cpu().numpy()
    assert np.abs(lo - golden["level/logits"]).max() < 1e-5
    assert np.array_equal(lo.argmax(-1), golden["level/logits"].argmax(-1))
    ls, _, _ = lvl.query(xg, order=0, mode="simt")
    assert np.abs(ls.cpu().numpy() - golden["level/logits"]).max() < 1e-5
    for n in (1, 127, 129, 300, 70001):
        g = torch.Generator().manual_seed(n)
        x = (torch.rand(n, 4, generator=g) * 2 - 1).cuda()
        a = lvl(x)
        b, _, _ = lvl.query(x, order=0, mode="simt")
        assert a.shape == (n, 8) and torch.isfinite(a).all()
        assert (a - b).abs().max().item() < 2e-5
        if n > 200:
            assert torch.equal(lvl(x[77:]), a[77:])
    torch.manual_seed(3)
    l3 = inf3dp.MLPClassifier(num_classes=3).cuda()
    x = (torch.rand(5000, 4) * 2 - 1).cuda()
    ref = l3.network(x)   # plain torch modules (cuBLAS FP32) as the reference of the same weights
    assert (l3(x) - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("engine", ENGINES_GRAD1)
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 127, 129, 257, 1000, 4097, 37889])
def test_ragged_sizes_and_batch_split_invariance(golden, engine, n):
    net = _net(golden)
    g = torch.Generator().manual_seed(n)
    x = (torch.rand(n, 3, generator=g) * 2 - 1).cuda()
    y, J, _ = net.query(x, order=1, mode=engine)
    assert y.shape == (n, 1) and J.shape == (n, 1, 3)
    if n == 0:
        return
    yo, Jo, _ = so.siren_jet(layers_from_golden(golden, "sdf"), x.cpu().numpy().astype(np.float64), order=1)
    assert np.abs(y.cpu().numpy() - yo).max() <= TOL_F
    assert so.angle_deg(J.cpu().numpy()[:, 0], Jo[:, 0]).max() <= TOL_ANGLE
    if n > 40:  # a point's result does not depend on its position in the batch
        k = 37
        y2, J2, _ = net.query(x[k:], order=1, mode=engine)
        assert torch.equal(y2, y[k:]) and torch.equal(J2, J[k:])


@pytest.mark.parametrize("engine", ENGINES_GRAD1)
def test_million_points_size_independent_properties(golden, engine):
    """Full-size property check (no oracle at this size): determinism, batch-split invariance, and agreement
    between engines on a sample checked against the oracle."""
    net = _net(golden)
    n = 1 << 20
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(n, 3, generator=g) * 2 - 1).cuda()
    y, J, _ = net.query(x, order=1, mode=engine)
    ya, Ja, _ = net.query(x, order=1, mode=engine)
    assert torch.equal(y, ya) and torch.equal(J, Ja)
    assert torch.isfinite(y).all() and torch.isfinite(J).all()
    idx = torch.randperm(n, generator=g)[:4096]
    yo, Jo, _ = so.siren_jet(layers_from_golden(golden, "sdf"), x[idx.cuda()].cpu().numpy().astype(np.float64), order=1)
    assert np.abs(y[idx.cuda()].cpu().numpy() - yo).max() <= TOL_F
    assert so.angle_deg(J[idx.cuda()].cpu().numpy()[:, 0], Jo[:, 0]).max() <= TOL_ANGLE


def test_weights_repack_after_update(golden):
    net = _net(golden)
    x = torch.from_numpy(golden["x"][:64]).cuda()
    y0, _, _ = net.query(x, order=0, mode="simt")
    with torch.no_grad():
        net.net.net[4][0].bias.add_(1.0)
    y1, _, _ = net.query(x, order=0, mode="simt")
    assert torch.allclose(y1, y0 + 1.0, atol=1e-6)


def test_reverse_engine_is_the_default_and_needs_its_workspace(golden):
    """AUTO routes value+gradient of single-output sine networks to the reverse-mode engine (bit-identical results),
    and the C ABI refuses that engine without the scratch buffer instead of silently running something else."""
    import ctypes
    import inf3dp
    from inf3dp import _lib
    net = _net(golden)
    x = torch.from_numpy(golden["x"]).cuda()
    ya, Ja, _ = net.query(x, order=1, mode="auto")
    yr, Jr, _ = net.query(x, order=1, mode="tc_reverse")
    assert torch.equal(ya, yr) and torch.equal(Ja, Jr)
    ws, bs = net._weights()
    blob = net._packed.get(ws, bs)
    y = torch.empty((len(x), 1), device="cuda"); J = torch.empty((len(x), 1, 3), device="cuda")
    L = _lib.lib()
    rc = L.inf3dp_siren_jet_ws(_lib.ptr(blob), _lib.ptr(x), len(x), 1, _lib.ptr(y), _lib.ptr(J), None,
                               _lib.MODE_TC_REVERSE, None, 0, _lib.stream_ptr(x.device))
    assert rc == _lib.EINVAL and b"workspace" in L.inf3dp_last_error()
    small = torch.empty(1024, dtype=torch.uint8, device="cuda")
    rc = L.inf3dp_siren_jet_ws(_lib.ptr(blob), _lib.ptr(x), len(x), 1, _lib.ptr(y), _lib.ptr(J), None,
                               _lib.MODE_TC_REVERSE, _lib.ptr(small), 1024, _lib.stream_ptr(x.device))
    assert rc == _lib.EINVAL
    # four-output networks are not served by the reverse engine
    q = _net(golden, "quat", 4)
    with pytest.raises(RuntimeError):
        q.query(x, order=1, mode="tc_reverse")


def test_reverse_engine_other_field_networks(golden):
    """A second and third weight set (different seeds: the slice / guidance field role) through the reverse engine."""
    import inf3dp
    for seed in (7, 99):
        torch.manual_seed(seed)
        net = inf3dp.SingleBVPNet(type="sine", in_features=3).cuda()
        layers = so.params_from_state_dict({k: v.detach().cpu().numpy() for k, v in net.state_dict().items()})
        g = torch.Generator().manual_seed(seed)
        x = (torch.rand(20000, 3, generator=g) * 2 - 1).cuda()
        y, J, _ = net.query(x, order=1, mode="tc_reverse")
        yo, Jo, _ = so.siren_jet(layers, x.cpu().numpy().astype(np.float64), order=1)
        assert np.abs(y.cpu().numpy() - yo).max() <= TOL_F
        assert so.angle_deg(J.cpu().numpy()[:, 0], Jo[:, 0]).max() <= TOL_ANGLE
        # gradient magnitude too (the angle alone would not see a wrong scale)
        rel = np.abs(np.linalg.norm(J.cpu().numpy()[:, 0], axis=1) / np.linalg.norm(Jo[:, 0], axis=1) - 1.0)
        assert rel.max() <= 2e-3


@pytest.mark.parametrize("engine", ENGINES_GRAD1)
def test_slice_field_vs_reference_golden(golden, engine):
    """The slice / guidance field network of the golden file (second weight set recorded from the reference)."""
    import inf3dp
    torch.manual_seed(2048)  # exp_scripts/train_sdf.py:13; construction order of tests/golden/make_golden.py: sdf, slice, ...
    inf3dp.SingleBVPNet(type="sine", in_features=3)
    net = inf3dp.SingleBVPNet(type="sine", in_features=3).cuda()   # the slice network (weights pinned by sha in test_dropin)
    x = torch.from_numpy(golden["x"]).cuda()
    y, J, _ = net.query(x, order=1, mode=engine)
    assert np.abs(y.cpu().numpy() - golden["slice/y"]).max() <= TOL_F
    assert so.angle_deg(J.cpu().numpy()[:, 0], golden["slice/grad"]).max() <= TOL_ANGLE


@pytest.mark.parametrize("n", [1, 11, 12, 13, 24, 25, 1000, 50021])
def test_order2_tensor_engine_ragged_sizes_vs_oracle(golden, n):
    """Tensor-core Hessian engine on ragged batch sizes (tiles are 12 points per CTA, 24 per CTA pair) against the
    fp64 oracle: value / gradient at the north_star tolerances, Hessian relative Frobenius <= 1e-3 (proposed tolerance,
    SURVEY.md 8c; measured ~3e-5), exact symmetry, batch-split invariance."""
    net = _net(golden)
    g = torch.Generator().manual_seed(100 + n)
    x = (torch.rand(n, 3, generator=g) * 2 - 1).cuda()
    y, J, H = net.query(x, order=2, mode="tc_precise")
    assert y.shape == (n, 1) and J.shape == (n, 1, 3) and H.shape == (n, 1, 3, 3)
    m = min(n, 4096)
    yo, Jo, Ho = so.siren_jet(layers_from_golden(golden, "sdf"), x[:m].cpu().numpy().astype(np.float64), order=2)
    assert np.abs(y[:m].cpu().numpy() - yo).max() <= TOL_F
    assert so.angle_deg(J[:m].cpu().numpy()[:, 0], Jo[:, 0]).max() <= TOL_ANGLE
    Hn = H.cpu().numpy()[:, 0]
    rel = np.linalg.norm((Hn[:m] - Ho[:, 0]).reshape(m, -1), axis=1) / np.linalg.norm(Ho[:, 0].reshape(m, -1), axis=1)
    assert rel.max() <= 1e-3
    assert np.array_equal(Hn, np.swapaxes(Hn, -1, -2))
    assert torch.isfinite(H).all()
    if n > 30:
        k = 17
        y2, J2, H2 = net.query(x[k:], order=2, mode="tc_precise")
        assert torch.equal(y2, y[k:]) and torch.equal(J2, J[k:]) and torch.equal(H2, H[k:])


@pytest.mark.parametrize("n", [1, 3, 31, 32, 33, 63, 64, 65, 1000, 50021])
def test_order2_forward_over_reverse_ragged_sizes_vs_oracle(golden, n):
    """Forward-over-reverse Hessian engine (mode tc_reverse with order 2: tiles are 32 points per CTA, 64 per CTA pair;
    diff_operators.py:6-44) on ragged batch sizes against the fp64 oracle: value / gradient at the north_star
    tolerances, Hessian relative Frobenius <= 1e-4 (measured ~4e-5 max), exact symmetry, batch-split invariance, and
    agreement with the second-order forward jet.  AUTO selects this engine (bit-identical results)."""
    net = _net(golden)
    g = torch.Generator().manual_seed(300 + n)
    x = (torch.rand(n, 3, generator=g) * 2 - 1).cuda()
    y, J, H = net.query(x, order=2, mode="tc_reverse")
    assert y.shape == (n, 1) and J.shape == (n, 1, 3) and H.shape == (n, 1, 3, 3)
    ya, Ja, Ha = net.query(x, order=2)
    assert torch.equal(ya, y) and torch.equal(Ja, J) and torch.equal(Ha, H)
    m = min(n, 4096)
    yo, Jo, Ho = so.siren_jet(layers_from_golden(golden, "sdf"), x[:m].cpu().numpy().astype(np.float64), order=2)
    assert np.abs(y[:m].cpu().numpy() - yo).max() <= TOL_F
    assert so.angle_deg(J[:m].cpu().numpy()[:, 0], Jo[:, 0]).max() <= TOL_ANGLE
    Hn = H.cpu().numpy()[:, 0]
    rel = np.linalg.norm((Hn[:m] - Ho[:, 0]).reshape(m, -1), axis=1) / np.linalg.norm(Ho[:, 0].reshape(m, -1), axis=1)
    assert rel.max() <= 1e-4
    assert np.array_equal(Hn, np.swapaxes(Hn, -1, -2))
    assert torch.isfinite(H).all()
    yj, Jj, Hj = net.query(x, order=2, mode="tc_precise")
    scale = float(Hj.abs().max())
    assert float((H - Hj).abs().max()) <= 2e-4 * scale and float((y - yj).abs().max()) <= 2e-6
    if n > 40:
        k = 17
        y2, J2, H2 = net.query(x[k:], order=2, mode="tc_reverse")
        assert torch.equal(y2, y[k:]) and torch.equal(J2, J[k:]) and torch.equal(H2, H[k:])


def test_order2_forward_over_reverse_value_gradient_match_the_reverse_engine(golden):
    """The primal row of the forward-over-reverse sweep runs the reverse engine's arithmetic: its value and gradient agree
    with an order-1 query to rounding noise, on the golden points and against the reference's golden outputs."""
    net = _net(golden)
    x = torch.from_numpy(golden["x"]).cuda()
    y2, J2, _ = net.query(x, order=2, mode="tc_reverse")
    y1, J1, _ = net.query(x, order=1, mode="tc_reverse")
    assert float((y2 - y1).abs().max()) <= 1e-6
    assert so.angle_deg(J2.cpu().numpy()[:, 0], J1.cpu().numpy()[:, 0]).max() <= 0.01
    assert np.abs(y2.cpu().numpy() - golden["sdf/y"]).max() <= TOL_F
    assert so.angle_deg(J2.cpu().numpy()[:, 0], golden["sdf/grad"]).max() <= TOL_ANGLE


def test_order2_forward_over_reverse_needs_the_workspace(golden):
    """INF3DP_MODE_TC_REVERSE with order 2 is refused without the scratch buffer; AUTO without one runs the forward jet."""
    from inf3dp import _lib
    net = _net(golden)
    x = torch.from_numpy(golden["x"]).cuda()
    ws, bs = net._weights()
    blob = net._packed.get(ws, bs)
    n = len(x)
    y = torch.empty((n, 1), device="cuda"); J = torch.empty((n, 1, 3), device="cuda"); H = torch.empty((n, 1, 3, 3), device="cuda")
    L = _lib.lib()
    rc = L.inf3dp_siren_jet_ws(_lib.ptr(blob), _lib.ptr(x), n, 2, _lib.ptr(y), _lib.ptr(J), _lib.ptr(H),
                               _lib.MODE_TC_REVERSE, None, 0, _lib.stream_ptr(x.device))
    assert rc == _lib.EINVAL and b"workspace" in L.inf3dp_last_error()
    rc = L.inf3dp_siren_jet_ws(_lib.ptr(blob), _lib.ptr(x), n, 2, _lib.ptr(y), _lib.ptr(J), _lib.ptr(H),
                               _lib.MODE_AUTO, None, 0, _lib.stream_ptr(x.device))
    assert rc == 0
    _, _, Hj = net.query(x, order=2, mode="tc_precise")
    assert torch.equal(H, Hj)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 255, 256, 257, 1000, 100003])
def test_value_only_on_the_cta_pair_engine(golden, n):
    """Order-0 queries of single-output sine fields run the forward sweep of the reverse engine alone (modules.py:143-160;
    what utils.field_querying and the SDF sweep of the collision query ask for): ragged sizes against the fp64 oracle and the
    golden values, bit-identical to the value an order-1 query returns, batch-split invariant; tc_reverse selects it too."""
    net = _net(golden)
    g = torch.Generator().manual_seed(500 + n)
    x = (torch.rand(n, 3, generator=g) * 2 - 1).cuda()
    y, J, H = net.query(x, order=0)
    assert y.shape == (n, 1) and J is None and H is None and torch.isfinite(y).all()
    y1, _, _ = net.query(x, order=1)
    assert torch.equal(y, y1)
    yr, _, _ = net.query(x, order=0, mode="tc_reverse")
    assert torch.equal(y, yr)
    m = min(n, 4096)
    yo, _, _ = so.siren_jet(layers_from_golden(golden, "sdf"), x[:m].cpu().numpy().astype(np.float64), order=0)
    assert np.abs(y[:m].cpu().numpy() - yo).max() <= 2e-6
    yt, _, _ = net.query(x, order=0, mode="tc_precise")
    assert float((y - yt).abs().max()) <= 1e-6
    if n > 40:
        y2, _, _ = net.query(x[17:], order=0)
        assert torch.equal(y2, y[17:])
    xg = torch.from_numpy(golden["x"]).cuda()
    yg, _, _ = net.query(xg, order=0)
    assert np.abs(yg.cpu().numpy() - golden["sdf/y"]).max() <= TOL_F


@pytest.mark.parametrize("gain", [1.0, 3.0, 8.0])
def test_reverse_engine_range_scale_with_large_weights(golden, gain):
    """Hidden weights scaled up (a stand-in for trained fields with high-frequency content; pre-activations reach
    tens of radians).  The backward sweep's range scale keeps every FP16 operand finite.  Accuracy: at gain 1 the
    north_star tolerances hold; at larger gains FP32 arithmetic itself (the SIMT engine = the reference's accuracy class)
    drifts from fp64 (0.07 deg at gain 3, 0.35 deg at gain 8), and the tensor engine stays within 4x of the FP32 engine."""
    import inf3dp
    torch.manual_seed(5)
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    with torch.no_grad():
        for i in (1, 2, 3):
            net.net.net[i][0].weight.mul_(gain)
    net.cuda()
    layers = so.params_from_state_dict({k: v.detach().cpu().numpy() for k, v in net.state_dict().items()})
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(8192, 3, generator=g) * 2 - 1).cuda()
    yo, Jo, _ = so.siren_jet(layers, x.cpu().numpy().astype(np.float64), order=1)
    err = {}
    for eng in ("simt", "tc_reverse"):
        y, J, _ = net.query(x, order=1, mode=eng)
        assert torch.isfinite(y).all() and torch.isfinite(J).all()
        err[eng] = (np.abs(y.cpu().numpy() - yo).max(), so.angle_deg(J.cpu().numpy()[:, 0], Jo[:, 0]).max())
    assert err["tc_reverse"][0] <= max(TOL_F, 4 * err["simt"][0])
    assert err["tc_reverse"][1] <= max(TOL_ANGLE, 4 * err["simt"][1])


@pytest.mark.parametrize("gain", [1.0, 3.0, 8.0])
def test_forward_over_reverse_hessian_with_large_weights(golden, gain):
    """The order-2 engine under the same stress: hidden weights scaled up (|H| reaches 1e6 at gain 8).  The tangent rows of the
    backward sweep ride on the reverse engine's range scale: every output stays finite (the second-order forward jet overflows
    FP16 at gain 8) and the Hessian error stays within 4x of the FP32 engine's own distance to fp64."""
    import inf3dp
    torch.manual_seed(5)
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    with torch.no_grad():
        for i in (1, 2, 3):
            net.net.net[i][0].weight.mul_(gain)
    net.cuda()
    layers = so.params_from_state_dict({k: v.detach().cpu().numpy() for k, v in net.state_dict().items()})
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(4096, 3, generator=g) * 2 - 1).cuda()
    _, _, Ho = so.siren_jet(layers, x.cpu().numpy().astype(np.float64), order=2)
    rel = {}
    for eng in ("simt", "tc_reverse"):
        y, J, H = net.query(x, order=2, mode=eng)
        assert torch.isfinite(y).all() and torch.isfinite(J).all() and torch.isfinite(H).all()
        Hn = H.cpu().numpy()[:, 0]
        rel[eng] = (np.linalg.norm((Hn - Ho[:, 0]).reshape(len(Hn), -1), axis=1) / np.linalg.norm(Ho[:, 0].reshape(len(Hn), -1), axis=1)).max()
    assert rel["tc_reverse"] <= max(1e-4, 4 * rel["simt"])


@pytest.mark.gpu
def test_packed_rows_equal_value_and_gradient(golden):
    """inf3dp_siren_value_grad_rows / inf3dp_unit_normals_rows: the packed (value, grad) rows are bit-identical to the
    separate y / J outputs of the same engine, also for ragged sizes and when written into a caller's buffer."""
    import inf3dp
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    net.load_state_dict({k[4:]: torch.from_numpy(np.asarray(golden[k])) for k in golden.files if k.startswith("sdf/net.")})
    net = net.cuda()
    g = torch.Generator(device="cpu").manual_seed(3)
    for n in (1, 127, 1024, 100_003):
        x = (torch.rand((n, 3), generator=g) * 2 - 1).cuda()
        y, J, _ = net.query(x, order=1, mode="tc_reverse")
        rows = net.query_rows(x)
        assert rows.shape == (n, 4)
        assert torch.equal(rows[:, :1], y) and torch.equal(rows[:, 1:], J[:, 0, :])
        buf = torch.full((n + 5, 4), -7.0, device="cuda")
        out = net.query_rows(x, out=buf[:n])
        assert out.data_ptr() == buf.data_ptr() and torch.equal(out, rows) and bool((buf[n:] == -7.0).all())
        assert torch.equal(inf3dp.unit_normals_rows(rows), inf3dp.unit_normals(J[:, 0, :].contiguous()))
    with pytest.raises(RuntimeError):
        net.query_rows(x.cpu())
    quat = inf3dp.SingleBVPNet(type="sine", in_features=3, out_features=4).cuda()
    with pytest.raises(RuntimeError):
        quat.query_rows(x)          # rows are the single-output field layout


@pytest.mark.gpu
@pytest.mark.parametrize("gain0", [1.0, 4.0, 12.0])
def test_large_first_layer_weights_sin_without_range_reduction(golden, gain0):
    """layer 0 evaluates sin / cos with MUFU (`sin.approx`, range reduction = one FP32 multiply by 1/2pi) and was
    only exercised at random-init first-layer weights (|30 z| <= ~42 rad).  Here W0 is scaled up so that the layer-0 arguments
    reach ~170 and ~500 rad (trained SIRENs with high-frequency first layers).  The argument itself is an FP32 quantity: at
    |z| = 500 rad one ulp is 3e-5 rad, which the reference's torch.sin sees too -- so the yardstick is the FP32 SIMT engine
    (sincosf with full range reduction = the reference's accuracy class), not a fixed tolerance: every tensor-core engine stays
    within 2x of its error against the fp64 oracle (plus the north-star floor), and everything is finite."""
    import inf3dp
    torch.manual_seed(7)
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    with torch.no_grad():
        net.net.net[0][0].weight.mul_(gain0)
        net.net.net[0][0].bias.mul_(gain0)
    net.cuda()
    layers = so.params_from_state_dict({k: v.detach().cpu().numpy() for k, v in net.state_dict().items()})
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(8192, 3, generator=g) * 2 - 1).cuda()
    xn = x.cpu().numpy().astype(np.float64)
    zmax = np.abs(30.0 * (xn @ layers[0][0].astype(np.float64).T + layers[0][1])).max()
    assert zmax > 25.0 * gain0
    yo, Jo, Ho = so.siren_jet(layers, xn, order=2)
    err = {}
    for eng in ("simt", "tc_precise", "tc_reverse"):
        y, J, _ = net.query(x, order=1, mode=eng)
        assert torch.isfinite(y).all() and torch.isfinite(J).all()
        err[eng] = (np.abs(y.cpu().numpy() - yo).max(), so.angle_deg(J.cpu().numpy()[:, 0], Jo[:, 0]).max())
    print(f"gain0 {gain0}: max |30 z0| {zmax:.0f} rad; (|df|, angle) " + ", ".join(f"{k} {v[0]:.1e}/{v[1]:.3f}" for k, v in err.items()))
    for eng in ("tc_precise", "tc_reverse"):
        assert err[eng][0] <= TOL_F + 2 * err["simt"][0], (eng, err)
        assert err[eng][1] <= TOL_ANGLE + 2 * err["simt"][1], (eng, err)
    # the second-order jet's layer 0 (sin, cos and their products with W0 rows) under the same arguments
    _, _, H = net.query(x[:2048], order=2)
    _, _, Hs = net.query(x[:2048], order=2, mode="simt")
    ref = Ho[:2048, 0].reshape(2048, -1)
    rel = lambda A: (np.linalg.norm(A.cpu().numpy()[:, 0].reshape(2048, -1) - ref, axis=1) / np.linalg.norm(ref, axis=1)).max()
    assert torch.isfinite(H).all() and rel(H) <= 1e-3 + 2 * rel(Hs), (rel(H), rel(Hs))


@pytest.mark.gpu
def test_principal_curvatures_vs_fp64_oracle_statistics(golden):
    """The curvature tail on 4096 points against the float64 restatement of diff_operators.py:167-220 fed with float64 jets
    (the r1 test compared 64 points with the reference's own FP32 result at 2e-2).  FP32 Hessians (rel. Frobenius ~1e-5) divided
    by |grad| set the floor; umbilic-like points (kappa_min ~ kappa_max) make the DIRECTIONS ill-conditioned, so directions are
    compared only where the two curvatures are separated."""
    import inf3dp
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    net.load_state_dict({k[4:]: torch.from_numpy(np.asarray(golden[k])) for k in golden.files if k.startswith("sdf/net.")})
    net = net.cuda()
    layers = [(golden[f"sdf/net.net.{i}.0.weight"], golden[f"sdf/net.net.{i}.0.bias"]) for i in range(5)]
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(4096, 3, generator=g) * 2 - 1)
    _, Jo, Ho = so.siren_jet(layers, x.numpy().astype(np.float64), order=2)
    ko, do = so.principal_curvatures(Jo[:, 0], Ho[:, 0])
    _, J, H = net.query(x.cuda(), order=2)
    k, d = inf3dp.principal_curvatures(J[:, 0], H[:, 0])
    k, d = k.cpu().numpy().astype(np.float64), d.cpu().numpy().astype(np.float64)
    scale = np.maximum(np.abs(ko).max(axis=1), 1.0)
    rel = np.abs(k - ko).max(axis=1) / scale
    print(f"curvatures: rel err median {np.median(rel):.1e} p99 {np.quantile(rel, 0.99):.1e} max {rel.max():.1e}")
    # measured on the B200: median 5.1e-6, p99 3.3e-5, max 1.4e-4
    assert np.median(rel) <= 5e-5 and np.quantile(rel, 0.99) <= 5e-4 and rel.max() <= 2e-3
    assert (k[:, 0] <= k[:, 1] + 1e-6).all()
    sep = (ko[:, 1] - ko[:, 0]) > 0.05 * scale
    cosang = np.abs((d[sep] * do[sep]).sum(-1))              # eigenvectors are defined up to sign
    assert sep.sum() > 1000 and np.quantile(cosang, 0.01) >= np.cos(np.radians(1.0))
    nrm = Jo[:, 0] / np.linalg.norm(Jo[:, 0], axis=1, keepdims=True)
    assert np.abs((d * nrm[:, None, :]).sum(-1)).max() <= 1e-3    # principal directions lie in the tangent plane
