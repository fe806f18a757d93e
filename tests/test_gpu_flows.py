"""GPU parity tests for the device-resident composite flows against golden
vectors recorded from the REFERENCE's own functions (tests/golden/make_golden_flows.py runs
toolpath_utils.batch_isosurface_projection and level_collision.CollTVSDF on torch CPU in the build container)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def flows():
    return np.load(os.path.join(GOLDEN, "flows_golden.npz"))


@pytest.fixture(scope="module")
def nets(golden):
    import inf3dp
    torch.manual_seed(2048)  # construction order of the golden scripts: sdf, slice, ... (weights pinned by SHA in test_dropin)
    sdf = inf3dp.SingleBVPNet(type="sine", in_features=3)
    slc = inf3dp.SingleBVPNet(type="sine", in_features=3)
    lvl = inf3dp.MLPClassifier(num_classes=8)
    lvl.load_state_dict({k[len("level/"):]: torch.from_numpy(np.asarray(golden[k])) for k in golden.files
                         if k.startswith("level/network")})
    return sdf.cuda(), slc.cuda(), lvl.cuda()


def test_admm_isosurface_projection_vs_reference(flows, nets):
    """toolpath_utils.py:275-372 incl. its quirk (z1 / z2 are never reassigned, so the x-step averages the original samples
    with the duals).  182 dependent field evaluations per sample; samples start 2e-3 from the target curve, the regime in
    which the reference's own fp32 iteration and an fp64 run of the same dynamics agree to 4e-7."""
    import inf3dp
    sdf, slc, _ = nets
    samples = torch.from_numpy(flows["proj/samples"]).cuda()
    meta = [(float(a), float(b), int(c)) for a, b, c in flows["proj/metadata"]]
    out = inf3dp.batch_isosurface_projection(samples, sdf, slc, meta).cpu().numpy()
    ref = flows["proj/result"]
    d = np.linalg.norm(out - ref, axis=1)
    moved = np.linalg.norm(ref - flows["proj/samples"], axis=1)
    print(f"projection: |dx| median {np.median(d):.2e} p99 {np.quantile(d, 0.99):.2e} max {d.max():.2e} (moved up to {moved.max():.3f})")
    assert np.median(d) <= 1e-5
    assert np.quantile(d, 0.99) <= 1e-4
    assert d.max() <= 1e-3
    # dict front end (toolpath_utils.py:234-272): short sub-contours are dropped, order preserved
    contours = {0.25: [flows["proj/samples"][:40], flows["proj/samples"][40:43]], -0.5: [flows["proj/samples"][43:90]]}
    re = inf3dp.iter_project_contours_batch(contours, sdf, slc)
    assert list(re.keys()) == [0.25, -0.5] and len(re[0.25]) == 1 and re[0.25][0].shape == (40, 3) and re[-0.5][0].shape == (47, 3)


def test_tvsdf_collision_vs_reference(flows, nets):
    """level_collision.py:73-255: identical collision mask up to points that sit on a decision boundary (sdf = 0,
    slice = maxslice, level arg-max ties) within the engines' ~1e-6 field difference; depths and gradient directions
    agree where both flag the point."""
    import inf3dp
    import copy
    sdf, slc, lvl = nets
    sdf = copy.deepcopy(sdf)
    with torch.no_grad():   # the golden run shifts the SDF output bias so that half of the cube is inside the part
        sdf.net.net[4][0].bias -= float(flows["coll/sdf_bias_shift"])
    nozzle = torch.from_numpy(flows["coll/nozzle"]).cuda().requires_grad_(True)
    levels = torch.from_numpy(flows["coll/levels"]).cuda()
    maxsl = torch.from_numpy(flows["coll/maxslices"]).cuda()
    depths, mask, wmask = inf3dp.collision_evaluation(nozzle, sdf, slc, lvl, 8, maxsl, levels, float(flows["coll/base_th"]))
    depths.sum().backward()
    d, m, g = depths.detach().cpu().numpy(), mask.cpu().numpy(), nozzle.grad.cpu().numpy()
    rd, rm, rg = flows["coll/depths"], flows["coll/mask"], flows["coll/grad"]
    assert m.shape == rm.shape == (96, 128) and wmask.shape == (96,)
    mismatch = (m != rm).sum()
    print(f"collision: {int(rm.sum())} reference hits, {int(m.sum())} ours, {int(mismatch)} mask mismatches; "
          f"max |d depth| on common hits {np.abs(d - rd)[m & rm].max():.2e}")
    assert mismatch <= 0.002 * m.size
    both = m & rm
    assert np.abs(d - rd)[both].max() <= 1e-4 or np.quantile(np.abs(d - rd)[both], 0.995) <= 1e-5
    assert (d[~m] == 0).all() and (d >= 0).all()
    # backward: -normalised saved direction (level_collision.py:239-255); compare directions on common hits
    gn, rgn = np.linalg.norm(g, axis=-1), np.linalg.norm(rg, axis=-1)
    hit = both & (rgn > 0)
    cosang = (g[hit] * rg[hit]).sum(-1) / np.maximum(gn[hit] * rgn[hit], 1e-30)
    assert np.quantile(cosang, 0.01) >= np.cos(np.radians(0.5))
    assert np.abs(gn[hit] - 1.0).max() <= 1e-5 and (gn[~m] == 0).all()
    assert np.array_equal(wmask.cpu().numpy(), m.any(-1))


def test_fused_collision_sweep_vs_per_point_composition(golden, nets):
    """inf3dp.quat_collision_sweep (nozzle posed inside the first stage kernel, results reduced per waypoint inside the
    stage kernels, nothing of size [N,M] materialised) against the composition a caller of the reference would write:
    utils.transform_nozzle_pcd_with_quaternion -> CollTVSDF per point -> max / any per waypoint.  The posed points differ in
    the last bit (torch's cross / normalize kernels vs the fused pose), so points that sit on a decision boundary may flip:
    waypoint masks agree on all but a handful of waypoints, depths agree where both flag the waypoint.  Also: cfg 4's real
    nozzle (data/rect_nozzle_shell.xyz[::10] / 150, 6 419 points, tests/golden/cfg1_golden.npz) on a waypoint slice."""
    import copy
    import math
    import inf3dp
    sdf, slc, lvl = nets
    sdf = copy.deepcopy(sdf)
    with torch.no_grad():
        sdf.net.net[4][0].bias -= 0.06
    torch.manual_seed(2048)
    quat = [inf3dp.SingleBVPNet(type="sine", in_features=3, out_features=o) for o in (1, 1, 1, 1, 4)][4].cuda()
    g = torch.Generator().manual_seed(11)
    for n, nozzle in ((20_000, None), (600, "shell")):
        u, v = torch.rand(n, generator=g) * 2 * math.pi, torch.rand(n, generator=g) * 2 * math.pi
        way = torch.stack([(0.6 + 0.3 * torch.cos(v)) * torch.cos(u), (0.6 + 0.3 * torch.cos(v)) * torch.sin(u), 0.3 * torch.sin(v)], 1).cuda()
        levels = ((way[:, 2] + 0.3) / 0.6 * 8).clamp(0, 7).long()
        maxsl = (torch.rand((n, 8), generator=g) * 0.1 - 0.05).cuda()
        if nozzle is None:
            t = torch.linspace(0, 1, 64)
            noz = (torch.stack([0.5 * t * torch.cos(t * 40), 0.5 * t * torch.sin(t * 40), 0.2 + 4.0 * t], 1) / 150.0 * 20.0).cuda()
        else:
            noz = torch.from_numpy(np.load(os.path.join(GOLDEN, "cfg1_golden.npz"))["nozzle_shell_every10_scaled"]).cuda()
            assert noz.shape == (6419, 3)
        d1, h1 = inf3dp.quat_collision_sweep(way, levels, maxsl, noz, quat, sdf, slc, lvl, 8, -0.9, waypoint_chunk=7001)
        d2, h2 = inf3dp.quat_collision_sweep_unfused(way, levels, maxsl, noz, quat, sdf, slc, lvl, 8, -0.9, waypoint_chunk=4096)
        flips = int((h1 != h2).sum())
        both = h1 & h2
        dd = (d1 - d2).abs()[both]
        print(f"collision sweep n={n} M={noz.shape[0]}: {int(h1.sum())} / {int(h2.sum())} colliding waypoints, {flips} flips, "
              f"max |d depth| {float(dd.max()) if dd.numel() else 0:.2e}")
        assert h1.any() and (~h1).any()
        assert flips <= max(2, 0.001 * n)
        assert dd.numel() == 0 or float(dd.quantile(0.999)) <= 1e-5
        assert (d1[~h1] == 0).all() and (d1 >= 0).all()


def test_collision_rejects_out_of_range_levels(nets):
    import inf3dp
    sdf, slc, lvl = nets
    pcd = torch.zeros((4, 8, 3), device="cuda")
    maxsl = torch.zeros((4, 8), device="cuda")
    with pytest.raises(RuntimeError, match="current_levels"):
        inf3dp.collision_evaluation(pcd, sdf, slc, lvl, 8, maxsl, torch.tensor([0, 1, 8, 2], device="cuda"), -0.9)
    with pytest.raises(RuntimeError, match="num_classes"):
        inf3dp.collision_evaluation(pcd, sdf, slc, lvl, 8, maxsl[:, :5], torch.tensor([0, 1, 2, 3], device="cuda"), -0.9)
    d, m, w = inf3dp.collision_evaluation(torch.full((4, 8, 3), 5.0, device="cuda"), sdf, slc, lvl, 8, maxsl,
                                          torch.tensor([0, 1, 2, 3], device="cuda"), -0.9)      # everything outside the cube
    assert d.shape == (4, 8) and not m.any() and not w.any() and (d == 0).all()
