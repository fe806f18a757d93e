"""The reference's OWN L3 helpers, unmodified, running on the drop-in on the B200 ("the unmodified
reference L3-L5 code must run against either").

oracle/build_ref.py stages byte copies of the reference's utils.py / modules.py / diff_operators.py / torchmeta and its
input point clouds under oracle/_ref/py (git-ignored, travels with the gpurun snapshot).  Here the SAME function objects
-- utils.field_query_with_grads (utils.py:217-266), utils.field_querying (:140-176), utils.quat_field_querying
(:178-214), utils.level_querying_with_confidence (:269-303) -- are called twice: once on the reference's
modules.SingleBVPNet on torch CUDA (the GPU oracle), once on inf3dp.SingleBVPNet holding the same state_dict, on
BASELINE.json configs[0] at its stated input: the full 16 018-point sdf_data/bunnyz.xyz cloud (and fertilityz.xyz).
Tolerances are the north star's: |df| <= 1e-4 of the bbox diagonal (3.46e-4), gradient angle <= 0.1 degrees.

tests/golden/cfg1_golden.npz holds what the reference produced for the same inputs on torch CPU in the build container
(tests/golden/make_golden_cfg1.py); the drop-in is compared with it as well, so this file still gates parity at the full
cfg-1 size on a box without oracle/_ref.  If NEITHER is available the tests FAIL (they never skip)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import build_ref
from oracle import siren_oracle as so

pytestmark = pytest.mark.gpu

TOL_F = 1e-4 * np.sqrt(12.0)
TOL_ANG = 0.1
CFG1 = os.path.join(GOLDEN, "cfg1_golden.npz")


class Decoder(torch.nn.Module):
    """exp_scripts/toolpath.py:21-30 (NNDecoder): the calling convention of every reference L4 function."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, coords):
        return self.model({"coords": coords})


@pytest.fixture(scope="module")
def cfg1():
    assert os.path.exists(CFG1), "tests/golden/cfg1_golden.npz is missing (tests/golden/make_golden_cfg1.py)"
    return np.load(CFG1)


@pytest.fixture(scope="module")
def ref():
    """(modules, diff_operators, utils) of the staged reference, or None when oracle/_ref/py did not travel."""
    if not build_ref.python_staged():
        yield None
        return
    keep = {k: sys.modules.get(k) for k in ("modules", "diff_operators", "utils")}
    mods = build_ref.load_python()
    yield mods
    for k, v in keep.items():          # leave no reference module behind for other test files
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def _seeded_reference_nets(modules):
    torch.manual_seed(2048)            # exp_scripts/train_sdf.py:13; construction order of tests/golden/make_golden.py
    with contextlib.redirect_stdout(io.StringIO()):
        return {name: modules.SingleBVPNet(type="sine", in_features=3, out_features=o)
                for name, o in (("sdf", 1), ("slice", 1), ("lattice1", 1), ("lattice2", 1), ("quat", 4))}


def _dropin_of(ref_net, out_features=1):
    import inf3dp
    net = inf3dp.SingleBVPNet(type="sine", in_features=3, out_features=out_features)
    net.load_state_dict(ref_net.state_dict())      # the reference's own keys, strict
    return net.cuda()


def _check(f, g, f_ref, g_ref, what):
    f, g, f_ref, g_ref = [np.asarray(t.detach().cpu() if torch.is_tensor(t) else t) for t in (f, g, f_ref, g_ref)]
    df = np.abs(f - f_ref).max()
    ang = so.angle_deg(g, g_ref).max()
    assert df <= TOL_F and ang <= TOL_ANG, (what, df, ang)
    return df, ang


@pytest.mark.parametrize("cloud,netname", [("bunnyz", "sdf"), ("fertilityz", "slice")])
def test_unmodified_field_query_with_grads_on_dropin_full_cloud(ref, cfg1, cloud, netname):
    coords = torch.from_numpy(cfg1[f"{cloud}/coords"])
    n = coords.shape[0]
    assert n == {"bunnyz": 16018, "fertilityz": 18634}[cloud]
    if ref is not None:
        modules, diff_operators, utils = ref
        nets = _seeded_reference_nets(modules)
        ref_net = nets[netname].cuda()
        ours = _dropin_of(ref_net)
        # the same unmodified function object on both networks (chunk loop, .cuda() hops, diff_operators.gradient)
        _, f_ref, g_ref = utils.field_query_with_grads(Decoder(ref_net), coords.clone(), no_print=True)
        _, f, g = utils.field_query_with_grads(Decoder(ours), coords.clone(), no_print=True)
        assert f.shape == (n,) and g.shape == (n, 3) and f.device.type == "cpu"
        _check(f, g, f_ref, g_ref, "drop-in vs reference net, both through utils.field_query_with_grads on CUDA")
        # small chunks (several per call) and the device-resident variant of the same helper
        _, f3, g3 = utils.field_query_with_grads(Decoder(ours), coords.clone(), max_batch=4096, no_print=True, return_on_cuda=True)
        assert f3.is_cuda and torch.equal(f3.cpu(), f) and torch.equal(g3.cpu(), g)
        # the reference's torch-CUDA result agrees with its own torch-CPU result recorded in the build container
        _check(f_ref, g_ref, cfg1[f"{cloud}/fields"], cfg1[f"{cloud}/grads"], "reference CUDA vs reference CPU golden")
    else:
        import inf3dp
        torch.manual_seed(2048)
        mine = {name: inf3dp.SingleBVPNet(type="sine", in_features=3, out_features=o)
                for name, o in (("sdf", 1), ("slice", 1))}
        f, g = inf3dp.field_query_with_grads(mine[netname].cuda(), coords.cuda())
    _check(f, g, cfg1[f"{cloud}/fields"], cfg1[f"{cloud}/grads"], "drop-in vs reference CPU golden (full cloud)")


def test_golden_only_path_full_bunnyz(cfg1):
    """Same gate without oracle/_ref: our own initialiser (same seed => the reference's weights, tests/test_dropin.py) and
    our fused helper against the reference's recorded torch-CPU answer on all 16 018 + 18 634 points."""
    import inf3dp
    torch.manual_seed(2048)
    nets = [inf3dp.SingleBVPNet(type="sine", in_features=3).cuda() for _ in range(2)]
    for cloud, net in (("bunnyz", nets[0]), ("fertilityz", nets[1])):
        x = torch.from_numpy(cfg1[f"{cloud}/coords"]).cuda()
        for engine in ("tc_reverse", "tc_precise", "simt"):
            f, g = inf3dp.field_query_with_grads(net, x, mode=engine)
            _check(f, g, cfg1[f"{cloud}/fields"], cfg1[f"{cloud}/grads"], (cloud, engine))


def test_unmodified_value_helpers_on_dropin(ref, cfg1):
    """utils.field_querying / quat_field_querying / level_querying_with_confidence, unmodified, on the drop-ins."""
    if ref is None:
        pytest.fail("oracle/_ref/py is absent: run `python oracle/build_ref.py` in the build container before gpurun")
    import inf3dp
    modules, diff_operators, utils = ref
    nets = _seeded_reference_nets(modules)
    coords = torch.from_numpy(cfg1["bunnyz/coords"])
    sdf_ref = nets["sdf"].cuda()
    v_ref = utils.field_querying(Decoder(sdf_ref), coords.clone(), no_print=True)
    v = utils.field_querying(Decoder(_dropin_of(sdf_ref)), coords.clone(), max_batch=5000, no_print=True)
    assert v.shape == v_ref.shape == (coords.shape[0],) and (v - v_ref).abs().max() <= TOL_F
    quat_ref = nets["quat"].cuda()
    q_ref = utils.quat_field_querying(Decoder(quat_ref), coords.clone(), no_print=True)
    q = utils.quat_field_querying(Decoder(_dropin_of(quat_ref, 4)), coords.clone(), no_print=True)
    assert q.shape == (coords.shape[0], 4) and (q - q_ref).abs().max() <= TOL_F
    # level classifier (exp_scripts/train_levels.py:35-51) through utils.level_querying_with_confidence
    torch.manual_seed(7)
    level_ref = torch.nn.Sequential(torch.nn.Linear(4, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
                                    torch.nn.Linear(256, 8)).cuda()
    level = inf3dp.MLPClassifier(num_classes=8)
    level.load_state_dict({f"network.{k}": t for k, t in level_ref.state_dict().items()})
    level.cuda()
    x4 = torch.cat([coords, v_ref[:, None]], 1)
    lab_ref, conf_ref = utils.level_querying_with_confidence(level_ref, x4.clone(), 8, no_print=True)
    lab, conf = utils.level_querying_with_confidence(level, x4.clone(), 8, no_print=True)
    agree = (lab == lab_ref).float().mean().item()
    assert agree >= 0.999 and (conf - conf_ref).abs().max() <= 1e-4, (agree, (conf - conf_ref).abs().max())


def test_unmodified_diff_operators_on_dropin_outputs(ref, golden):
    """diff_operators.gradient / hessian3d / hessian / laplace / divergence of the REFERENCE module applied to the drop-in's
    outputs (autograd bridge incl. double backward), next to the same calls on the reference network."""
    if ref is None:
        pytest.fail("oracle/_ref/py is absent: run `python oracle/build_ref.py` in the build container before gpurun")
    modules, diff_operators, utils = ref
    nets = _seeded_reference_nets(modules)
    ref_net = nets["sdf"].cuda()
    ours = _dropin_of(ref_net)
    x = torch.from_numpy(golden["x"][:256]).cuda()
    o_r, o_m = ref_net({"coords": x}), ours({"coords": x})
    g_r = diff_operators.gradient(o_r["model_out"], o_r["model_in"])
    g_m = diff_operators.gradient(o_m["model_out"], o_m["model_in"])
    _check(o_m["model_out"][:, 0], g_m, o_r["model_out"][:, 0], g_r, "gradient")
    h_r = diff_operators.hessian3d(o_r["model_out"], o_r["model_in"])
    h_m = diff_operators.hessian3d(o_m["model_out"], o_m["model_in"])
    rel = ((h_m - h_r).flatten(1).norm(dim=1) / h_r.flatten(1).norm(dim=1)).max().item()
    assert h_m.shape == h_r.shape == (256, 1, 3, 3) and rel <= 1e-3, rel
    lap_r = diff_operators.laplace(o_r["model_out"], o_r["model_in"])
    lap_m = diff_operators.laplace(o_m["model_out"], o_m["model_in"])
    assert (lap_m - lap_r).abs().max() <= 1e-3 * lap_r.abs().max()
    ob_r, ob_m = ref_net({"coords": x[None]}), ours({"coords": x[None]})       # [1,N,3] DataLoader layout
    hb_r, st_r = diff_operators.hessian(ob_r["model_out"], ob_r["model_in"])
    hb_m, st_m = diff_operators.hessian(ob_m["model_out"], ob_m["model_in"])
    assert st_r == st_m == 0 and hb_m.shape == hb_r.shape
    assert ((hb_m - hb_r).flatten(2).norm(dim=2) / hb_r.flatten(2).norm(dim=2)).max().item() <= 1e-3


def test_patch_reference_rebinds_the_class(ref):
    if ref is None:
        pytest.fail("oracle/_ref/py is absent: run `python oracle/build_ref.py` in the build container before gpurun")
    import inf3dp
    modules, _, _ = ref
    original = modules.SingleBVPNet
    try:
        assert inf3dp.patch_reference() is True
        assert modules.SingleBVPNet is inf3dp.SingleBVPNet
        net = modules.SingleBVPNet(type="sine", in_features=3, out_features=1, final_layer_factor=1).cuda()
        out = net({"coords": torch.zeros(5, 3, device="cuda")})
        assert out["model_out"].shape == (5, 1) and out["model_in"].requires_grad
    finally:
        modules.SingleBVPNet = original


def test_unmodified_collision_evaluation_vs_fused_stages_real_nozzle(ref, cfg1):
    """BASELINE.json configs[3] at the reference's own nozzle: level_collision.collision_evaluation (level_collision.py:259-281,
    the UNMODIFIED reference function: CollTVSDF.forward + its utils helpers, staged in oracle/_ref/py) on the reference's
    networks on torch CUDA, against inf3dp.collision_evaluation (fused predicate / compaction stages, csrc/collision.cu) on the
    drop-in networks holding the same weights.  Nozzle = data/rect_nozzle_shell.xyz[::10] / 150 (6 419 points,
    level_collision.py:301,463), posed by the reference's own utils.transform_nozzle_pcd_with_quaternion at 48 waypoints
    (308 112 nozzle points), plus the backward pass (normalised saved directions)."""
    if ref is None:
        pytest.fail("oracle/_ref/py is absent: run `python oracle/build_ref.py` in the build container before gpurun")
    import copy
    import importlib
    import math
    import inf3dp
    modules, diff_operators, utils = ref
    level_collision = importlib.import_module("level_collision")
    nets = _seeded_reference_nets(modules)
    sdf_ref, slice_ref, quat_ref = nets["sdf"].cuda(), nets["slice"].cuda(), nets["quat"].cuda()
    with torch.no_grad():
        sdf_ref.net.net[4][0].bias -= 0.06       # the random-init SDF is positive everywhere: put half of the cube inside
    torch.manual_seed(7)
    level_ref = torch.nn.Sequential(torch.nn.Linear(4, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
                                    torch.nn.Linear(256, 8)).cuda()
    sdf, slc = _dropin_of(sdf_ref), _dropin_of(slice_ref)
    level = inf3dp.MLPClassifier(num_classes=8)
    level.load_state_dict({f"network.{k}": t for k, t in level_ref.state_dict().items()})
    level.cuda()
    nozzle = torch.from_numpy(cfg1["nozzle_shell_every10_scaled"]).cuda()
    g = torch.Generator().manual_seed(3)
    n = 48
    u, v = torch.rand(n, generator=g) * 2 * math.pi, torch.rand(n, generator=g) * 2 * math.pi
    way = torch.stack([(0.6 + 0.3 * torch.cos(v)) * torch.cos(u), (0.6 + 0.3 * torch.cos(v)) * torch.sin(u), 0.3 * torch.sin(v)], 1).cuda()
    levels = ((way[:, 2] + 0.3) / 0.6 * 8).clamp(0, 7).long()
    maxsl = (torch.rand((n, 8), generator=g) * 0.1 - 0.05).cuda()
    quats = utils.quat_field_querying(Decoder(quat_ref), way.cpu(), no_print=True, return_on_cuda=True)
    pcd = utils.transform_nozzle_pcd_with_quaternion(nozzle, way, quats).contiguous()
    assert pcd.shape == (n, 6419, 3)
    p_ref = pcd.clone().requires_grad_(True)
    d_ref, m_ref, w_ref = level_collision.collision_evaluation(p_ref, Decoder(sdf_ref), Decoder(slice_ref), level_ref, 8, maxsl, levels, -0.9)
    d_ref.sum().backward()
    p_our = pcd.clone().requires_grad_(True)
    d, m, w = inf3dp.collision_evaluation(p_our, sdf, slc, level, 8, maxsl, levels, -0.9)
    d.sum().backward()
    mism = int((m != m_ref).sum())
    both = m & m_ref
    dd = (d - d_ref).detach().abs()[both]
    print(f"collision (reference's own function on CUDA vs fused stages): {int(m_ref.sum())} / {int(m.sum())} hits of {m.numel()} points, "
          f"{mism} mask mismatches, max |d depth| on common hits {float(dd.max()):.2e}")
    assert m_ref.any() and (~m_ref).any()
    assert mism <= 2e-4 * m.numel()                        # points on a decision boundary within the engines' ~1e-6 field difference
    assert float(dd.quantile(0.999)) <= 1e-5 and float(dd.max()) <= 1e-3
    assert int((w != w_ref).sum()) <= 1
    gr, go = p_ref.grad[both], p_our.grad[both]
    has = gr.norm(dim=-1) > 0
    cosang = torch.nn.functional.cosine_similarity(gr[has], go[has], dim=-1)
    assert float(cosang.quantile(0.01)) >= math.cos(math.radians(0.5))
    # and the sweep (nozzle posed inside the first stage kernel, per-waypoint reduction inside the kernels)
    quat = _dropin_of(quat_ref, 4)
    dw, hw = inf3dp.quat_collision_sweep(way, levels, maxsl, nozzle, quat, sdf, slc, level, 8, -0.9)
    assert int((hw != w_ref).sum()) <= 1
    assert float((dw - d_ref.detach().max(dim=-1).values).abs()[hw & w_ref].max()) <= 1e-4
