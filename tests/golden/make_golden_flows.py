"""Generate tests/golden/flows_golden.npz by running the REFERENCE's own composite flows (torch CPU) in the build
container (it has /root/reference mounted; the GPU box has not -- the committed .npz is what travels):

  * toolpath_utils.batch_isosurface_projection (toolpath_utils.py:275-372): ADMM projection of contour samples onto
    {sdf = c1} /\\ {slice = c2}                                                    -> SURVEY.md section 8f, row N1
  * level_collision.CollTVSDF forward + backward (level_collision.py:73-255): TV-SDF collision depths, mask and the
    saved gradient directions                                                     -> SURVEY.md section 8f, row N3

The reference hard-codes `.cuda()` in its query helpers (utils.py:161,175,236,261; level_collision.py:205); on this
CPU-only container `Tensor.cuda` / `torch.cuda.empty_cache` are shimmed to no-ops, nothing else
of the reference is touched.  Networks: the random-init fixture nets of seed 2048 in the order of make_golden.py
(SDF, slice, lattice1, lattice2, quaternion, level MLP); the SDF net and the level MLP are the ones stored whole in
siren_golden.npz, the slice net is rebuilt from the seed by the tests (its SHA-256 is pinned there).

    python tests/golden/make_golden_flows.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


import importlib.abc
import importlib.machinery
import types


class _Anything:
    """Attribute-swallowing placeholder for the visualisation / IO packages the reference imports at module level."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


class StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Any sub-module of the absent third-party packages imports as an empty stub (none of them is on the hot path)."""
    TOPS = {"matplotlib", "pyvista", "pyvistaqt", "skimage", "plyfile", "trimesh", "potpourri3d", "configargparse", "qupkg",
            "h5py", "PyQt5", "vtk", "open3d", "cv2", "imageio", "seaborn", "networkx", "tensorboard", "glm", "cupy"}

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in self.TOPS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []
        m.__getattr__ = lambda attr: (_ for _ in ()).throw(AttributeError(attr)) if attr.startswith("__") else _Anything()
        m.__all__ = []
        return m

    def exec_module(self, module):
        pass


class Decoder(torch.nn.Module):
    """exp_scripts/toolpath.py:21-30 (NNDecoder) without the checkpoint loading."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, coords):
        return self.model({"coords": coords})


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.empty_cache = lambda: None
    modules, _ = import_reference()
    for name in [n for n in sys.modules if n.split(".")[0] in StubFinder.TOPS]:
        del sys.modules[name]    # the flat stubs of make_golden.py are not packages (`matplotlib.colors` would fail)
    sys.meta_path.insert(0, StubFinder())
    with contextlib.redirect_stdout(io.StringIO()):
        import toolpath_utils
        import level_collision
        torch.manual_seed(2048)
        nets = {}
        for name, outf in [("sdf", 1), ("slice", 1), ("lattice1", 1), ("lattice2", 1), ("quat", 4)]:
            nets[name] = modules.SingleBVPNet(type="sine", in_features=3, out_features=outf)
        level = torch.nn.Sequential(torch.nn.Linear(4, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
                                    torch.nn.Linear(256, 8))
    sdf_dec, slice_dec = Decoder(nets["sdf"]), Decoder(nets["slice"])
    out = {}

    # ---- N1: ADMM iso-surface projection ------------------------------------------------------------------------------
    rng = np.random.default_rng(11)
    contours, metadata = [], []
    for j in range(12):
        centre = rng.uniform(-0.7, 0.7, 3)
        n = int(rng.integers(20, 80))
        # 2e-3 spread: the regime in which the reference's own iteration is well conditioned (its fp32 and an fp64
        # evaluation of the same dynamics end within 4e-7 of each other; at 2e-2 they end up to 0.2 apart)
        pts = (centre[None, :] + 0.002 * rng.standard_normal((n, 3))).astype(np.float32)
        with torch.enable_grad():
            c = torch.from_numpy(centre[None, :].astype(np.float32))
            f1 = float(nets["sdf"]({"coords": c})["model_out"].detach().reshape(-1)[0])
            f2 = float(nets["slice"]({"coords": c})["model_out"].detach().reshape(-1)[0])
        contours.append(pts)
        metadata.append((f2, f1, n))       # (iso_value, sdf_value, subcontour_len), toolpath_utils.py:255
    samples = torch.from_numpy(np.vstack(contours)).float()
    with contextlib.redirect_stdout(io.StringIO()):
        proj = toolpath_utils.batch_isosurface_projection(samples.clone(), sdf_dec, slice_dec, metadata)
    out["proj/samples"] = samples.numpy()
    out["proj/metadata"] = np.array(metadata, dtype=np.float64)
    out["proj/result"] = proj.detach().numpy()
    with torch.enable_grad():
        _, f1, _ = toolpath_utils.field_query_with_grads(sdf_dec, proj.detach().clone(), no_print=True, return_on_cuda=True)
        _, f2, _ = toolpath_utils.field_query_with_grads(slice_dec, proj.detach().clone(), no_print=True, return_on_cuda=True)
    c1 = np.repeat([m[1] for m in metadata], [m[2] for m in metadata])
    c2 = np.repeat([m[0] for m in metadata], [m[2] for m in metadata])
    print("projection: residual |sdf - c1| max %.2e, |slice - c2| max %.2e, moved %.3e max" % (
        np.abs(f1.numpy() - c1).max(), np.abs(f2.numpy() - c2).max(),
        np.linalg.norm(proj.detach().numpy() - samples.numpy(), axis=1).max()))

    # ---- N3: TV-SDF collision ------------------------------------------------------------------------------------------------
    rng = np.random.default_rng(12)
    N, M, C = 96, 128, 8
    nozzle = torch.from_numpy(rng.uniform(-1.08, 1.08, (N, M, 3)).astype(np.float32)).requires_grad_(True)
    levels = torch.from_numpy(rng.integers(0, C, N).astype(np.int64))
    with torch.enable_grad():
        probe = torch.from_numpy(rng.uniform(-1, 1, (4096, 3)).astype(np.float32))
        sl = nets["slice"]({"coords": probe})["model_out"].detach().numpy().reshape(-1)
    maxslices = torch.from_numpy(rng.uniform(np.quantile(sl, 0.2), np.quantile(sl, 0.8), (N, C)).astype(np.float32))
    base_th = -0.9
    # the random-init SDF net is positive everywhere (0.001 .. 0.14): shift its output bias so that about half of the
    # cube is "inside the part" and the TV-SDF branches (level / slice predicates, Eq. 27, two-step push) are exercised
    sdf_shift = 0.06
    with torch.no_grad():
        nets["sdf"].net.net[4][0].bias -= sdf_shift
    out["coll/sdf_bias_shift"] = np.float32(sdf_shift)
    with contextlib.redirect_stdout(io.StringIO()):
        depths, mask = level_collision.CollTVSDF.apply(nozzle, sdf_dec, slice_dec, level, C, maxslices, levels, base_th)
        depths.sum().backward()
    out["coll/nozzle"] = nozzle.detach().numpy()
    out["coll/levels"] = levels.numpy()
    out["coll/maxslices"] = maxslices.numpy()
    out["coll/base_th"] = np.float32(base_th)
    out["coll/depths"] = depths.detach().numpy()
    out["coll/mask"] = mask.numpy()
    out["coll/grad"] = nozzle.grad.numpy()
    print("collision: %d of %d points in collision, %d with non-zero depth, max depth %.4f" % (
        int(mask.sum()), mask.numel(), int((depths > 0).sum()), float(depths.max())))
    np.savez_compressed(os.path.join(HERE, "flows_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "flows_golden.npz"))


if __name__ == "__main__":
    main()
