"""Generate tests/golden/siren_golden.npz by running the REFERENCE itself (torch CPU) in the build container.

Run from the repo root in the container that has /root/reference mounted:
    python tests/golden/make_golden.py
It cannot run on the GPU box (no /root/reference there); the committed .npz files are what travels.

What it records:
  * the random-init fixture nets of seed 2048 (exp_scripts/train_sdf.py:13), built in the order
    SDF, slice, lattice1, lattice2, quaternion(out=4), level MLP(8 classes); the SDF and quaternion nets and the
    level MLP are stored whole, the others as SHA-256 of their weights (our own initialiser must reproduce them)
  * reference outputs on 512 bunnyz points (normalised by sdf_meshing.PCD) and 512 uniform points:
    SingleBVPNet.forward, diff_operators.gradient, diff_operators.hessian3d / hessian,
    min_max_curvs_and_vectors on 64 points, MLPClassifier logits, gen_voxle_queries(8)
Also asserts that oracle/siren_oracle.py (fp64 jet) and oracle/siren_torch_port.py agree with the reference.
"""
import contextlib
import hashlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def import_reference():
    for name in ["h5py", "matplotlib", "matplotlib.pyplot", "pyvista", "pyvistaqt", "skimage", "skimage.measure",
                 "plyfile", "trimesh", "potpourri3d", "configargparse", "qupkg", "qupkg.extract_isocontours"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__getattr__ = lambda attr, _n=name: (_ for _ in ()).throw(AttributeError(attr)) if attr.startswith("__") else object
            sys.modules[name] = m
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "exp_scripts"))
    import modules  # noqa
    import diff_operators  # noqa
    return modules, diff_operators


def sha(sd):
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().numpy()).tobytes())
    return h.hexdigest()


def main():
    sys.path.insert(0, ROOT)
    modules, diff_operators = import_reference()
    from oracle import siren_oracle as so
    from oracle import siren_torch_port as port

    torch.manual_seed(2048)
    nets = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for name, outf in [("sdf", 1), ("slice", 1), ("lattice1", 1), ("lattice2", 1), ("quat", 4)]:
            nets[name] = modules.SingleBVPNet(type="sine", in_features=3, out_features=outf)
    # exp_scripts/train_levels.py:35-51 (MLPClassifier); rebuilt here because importing the script runs training
    level = torch.nn.Sequential(torch.nn.Linear(4, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
                                torch.nn.Linear(256, 8))

    out = {}
    for name in ["sdf", "quat"]:
        for k, v in nets[name].state_dict().items():
            out[f"{name}/{k}"] = v.numpy().copy()
    for i, idx in enumerate([0, 2, 4]):
        out[f"level/network.{idx}.weight"] = level[idx].weight.detach().numpy().copy()
        out[f"level/network.{idx}.bias"] = level[idx].bias.detach().numpy().copy()
    out["sha_names"] = np.array(list(nets.keys()))
    out["sha_values"] = np.array([sha(n.state_dict()) for n in nets.values()])

    # inputs
    pc = np.genfromtxt(os.path.join(REF, "sdf_data", "bunnyz.xyz"))
    coords = so.normalise_point_cloud(pc[:, :3])
    # check the normalisation restatement against the reference class
    import sdf_meshing
    with contextlib.redirect_stdout(io.StringIO()):
        ref_pcd = sdf_meshing.PCD(os.path.join(REF, "sdf_data", "bunnyz.xyz"))
    assert np.array_equal(ref_pcd.coords, coords), "PCD normalisation restatement differs"
    rng = np.random.default_rng(0)
    sel = rng.choice(len(coords), 512, replace=False)
    x_bunny = coords[sel].astype(np.float32)
    x_unif = (rng.random((512, 3)) * 2 - 1).astype(np.float32)
    X = np.concatenate([x_bunny, x_unif], 0)
    out["x"] = X
    out["bunny_sel"] = sel.astype(np.int32)

    for name in ["sdf", "slice", "quat"]:
        net = nets[name]
        o = net({"coords": torch.from_numpy(X)})
        y, xin = o["model_out"], o["model_in"]
        out[f"{name}/y"] = y.detach().numpy().copy()
        if name != "quat":
            g = diff_operators.gradient(y, xin)
            out[f"{name}/grad"] = g.detach().numpy().copy()
            h = diff_operators.hessian3d(y[:128], xin) if False else None
        # jacobian per channel via gradient with one-hot grad_outputs (diff_operators.py:128-132)
        J = []
        for c in range(y.shape[1]):
            go = torch.zeros_like(y)
            go[:, c] = 1
            J.append(diff_operators.gradient(y, xin, go).detach().numpy().copy())
        out[f"{name}/jac"] = np.stack(J, 1)
        # oracle (fp64 jet) and torch port agree with the reference
        layers = so.params_from_state_dict({k: v.numpy() for k, v in net.state_dict().items()})
        yo, Jo, _ = so.siren_jet(layers, X.astype(np.float64), order=1)
        assert np.abs(yo - out[f"{name}/y"]).max() < 2e-6, name
        assert np.abs(Jo - out[f"{name}/jac"]).max() < 5e-5, name
        pm = port.PortSiren([(W.astype(np.float32), b.astype(np.float32)) for W, b in layers])
        po = pm({"coords": torch.from_numpy(X)})
        assert torch.equal(po["model_out"], y), "torch port forward differs from reference"
        if name != "quat":
            pg = port.gradient(po["model_out"], po["model_in"])
            assert torch.equal(pg, g), "torch port gradient differs from reference"

    # Hessians on 128 points (sdf net): hessian3d and batched hessian
    net = nets["sdf"]
    xs = torch.from_numpy(X[448:576])  # 64 bunny + 64 uniform
    o = net({"coords": xs})
    h3 = diff_operators.hessian3d(o["model_out"], o["model_in"])
    out["sdf/hess_x"] = xs.numpy().copy()
    out["sdf/hess3d"] = h3.detach().numpy().copy()
    ob = net({"coords": xs[None]})
    hb, status = diff_operators.hessian(ob["model_out"], ob["model_in"])
    assert status == 0
    out["sdf/hess_batched"] = hb.detach().numpy().copy()
    layers = so.params_from_state_dict({k: v.numpy() for k, v in net.state_dict().items()})
    _, _, Ho = so.siren_jet(layers, xs.numpy().astype(np.float64), order=2)
    rel = np.linalg.norm((Ho - out["sdf/hess3d"]).reshape(len(xs), -1), axis=1) / np.linalg.norm(Ho.reshape(len(xs), -1), axis=1)
    assert rel.max() < 1e-3, rel.max()
    out["sdf/hess_rel_fro_fp32_vs_fp64_max"] = np.array(rel.max())

    # curvature tail on 64 points
    class Dec(torch.nn.Module):
        def forward(self, c):
            return net({"coords": c})
    kap, pdir = diff_operators.min_max_curvs_and_vectors(Dec(), xs[:64])
    out["sdf/curv_kappa"] = kap.detach().numpy().copy()
    out["sdf/curv_dirs"] = pdir.detach().numpy().copy()
    yo, Jo, Ho = so.siren_jet(layers, xs[:64].numpy().astype(np.float64), order=2)
    ko, do = so.principal_curvatures(Jo[:, 0], Ho[:, 0])
    assert np.abs(ko - out["sdf/curv_kappa"]).max() < 2e-2 * max(1.0, np.abs(ko).max()), np.abs(ko - out["sdf/curv_kappa"]).max()

    # level MLP logits on (xyz, slice value)
    sl = out["slice/y"][:, 0:1]
    lx = np.concatenate([X, sl], 1).astype(np.float32)
    out["level/x"] = lx
    out["level/logits"] = level(torch.from_numpy(lx)).detach().numpy().copy()
    lo = so.relu_mlp([(out["level/network.0.weight"], out["level/network.0.bias"]),
                      (out["level/network.2.weight"], out["level/network.2.bias"]),
                      (out["level/network.4.weight"], out["level/network.4.bias"])], lx)
    assert np.abs(lo - out["level/logits"]).max() < 1e-5

    # grid generator quirk
    g8, _, _ = sdf_meshing.gen_voxle_queries(8)
    out["grid8"] = g8.numpy().copy()
    assert np.array_equal(so.gen_voxel_queries(8), out["grid8"]), "gen_voxle_queries restatement differs"
    g33, _, _ = sdf_meshing.gen_voxle_queries(33)
    assert np.array_equal(so.gen_voxel_queries(33), g33.numpy())

    path = os.path.join(HERE, "siren_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
