"""Generate tests/golden/cfg1_golden.npz: BASELINE.json configs[0] at its STATED input -- the reference's own
utils.field_query_with_grads (utils.py:217-266, max_batch 32^3, torch CPU with Tensor.cuda shimmed to the identity)
on the FULL sdf_data/bunnyz.xyz point cloud (16 018 points, normalised by sdf_meshing.PCD, sdf_meshing.py:17-38) with
the seed-2048 random-init SDF network (exp_scripts/train_sdf.py:13), and the same for sdf_data/fertilityz.xyz
(18 634 points) with the slice network (second network built after the seed).

Also records the reference's own level_collision.CollTVSDF forward (level_collision.py:73-237) on a cfg-4 shaped
input whose nozzle cloud is the reference's data/rect_nozzle_shell.xyz[::10] (6 419 points, scaled 1/150,
configs/collision_config.yaml:118) posed at 8 waypoints.

Run in the build container (needs /root/reference, or the staged copy under oracle/_ref/py):
    python tests/golden/make_golden_cfg1.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    from oracle import build_ref
    from oracle import siren_oracle as so
    if not build_ref.python_staged():
        build_ref.stage_python()
    modules, diff_operators, utils = build_ref.load_python(cpu_shim=True)
    import sdf_meshing

    torch.manual_seed(2048)
    with contextlib.redirect_stdout(io.StringIO()):
        sdf = modules.SingleBVPNet(type="sine", in_features=3)
        slc = modules.SingleBVPNet(type="sine", in_features=3)

    class Dec(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.model = m

        def forward(self, c):
            return self.model({"coords": c})

    out = {}
    for name, rel, net in (("bunnyz", "sdf_data/bunnyz.xyz", sdf), ("fertilityz", "sdf_data/fertilityz.xyz", slc)):
        path = os.path.join(build_ref.py_dir(), rel)
        with contextlib.redirect_stdout(io.StringIO()):
            pcd = sdf_meshing.PCD(path)
        coords = np.asarray(pcd.coords)
        assert np.array_equal(coords, so.normalise_point_cloud(np.genfromtxt(path)[:, :3]))
        x = torch.from_numpy(coords).float()
        _, f, g = utils.field_query_with_grads(Dec(net), x.clone(), no_print=True)
        out[f"{name}/coords"] = x.numpy().copy()
        out[f"{name}/fields"] = f.numpy().copy()
        out[f"{name}/grads"] = g.numpy().copy()
        # float64 truth of the same weights (sets the floor under the tolerances)
        layers = so.params_from_state_dict({k: v.numpy() for k, v in net.state_dict().items()})
        yo, Jo, _ = so.siren_jet(layers, x.numpy().astype(np.float64), order=1)
        out[f"{name}/ref_vs_fp64_absmax"] = np.array(np.abs(yo[:, 0] - f.numpy()).max())
        out[f"{name}/ref_vs_fp64_angle_max_deg"] = np.array(so.angle_deg(g.numpy(), Jo[:, 0]).max())
        print(name, x.shape, float(out[f"{name}/ref_vs_fp64_absmax"]), float(out[f"{name}/ref_vs_fp64_angle_max_deg"]))
    noz = np.genfromtxt(os.path.join(build_ref.py_dir(), "data", "rect_nozzle_shell.xyz"))[::10, :3]
    out["nozzle_shell_every10_scaled"] = (noz / 150.0).astype(np.float32)    # level_collision.py:301,463 + collision_config.yaml:118
    print("nozzle", out["nozzle_shell_every10_scaled"].shape)
    path = os.path.join(HERE, "cfg1_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
