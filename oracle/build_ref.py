"""ORACLE build recipe (test infrastructure) -- compile the REFERENCE contour_cuda extension into oracle/_ref/.

The reference's native path is three source files (contour_cuda/{ext.cpp,isocontours.cu,intersection.cu});
they are compiled from where they lie under /root/reference.  Nothing from the reference is committed:
oracle/_ref/ is git-ignored (it still travels to the GPU box with the gpurun snapshot).

Two build-only adaptations are applied to a scratch copy under oracle/_ref/src (behaviour unchanged):
  * `X.type()` -> `X.scalar_type()` in the AT_DISPATCH_FLOATING_TYPES calls (isocontours.cu:242,
    intersection.cu:258): torch >= 2.x no longer converts DeprecatedTypeProperties to ScalarType
  * an empty glm/glm.hpp: includes/utils.h:2 includes it, nothing uses it, and the reference's
    third_party/glm is not vendored (setup.py:17)
The module is used by tests/ (parity on the GPU), tests/golden/make_contour_golden.py (records goldens on the
B200) and bench.py's GPU-comparator leg only.

stage_python() additionally stages the reference's PYTHON query path UNMODIFIED (byte copies) into oracle/_ref/py:
utils.py, modules.py, diff_operators.py, level_collision.py, sdf_meshing.py, toolpath_utils.py, vis.py, the vendored
torchmeta package, and the three input point clouds the BASELINE configs name (sdf_data/bunnyz.xyz,
sdf_data/fertilityz.xyz, data/rect_nozzle_shell.xyz).  A Python reference cannot otherwise travel to the GPU box;
staged this way the unmodified L3 helpers (utils.field_query_with_grads ...) run there on torch CUDA (the GPU oracle
of SURVEY.md section 8c) and on the host cores (bench.py --impl reference, kind "reference").  load_python() imports
them with stand-ins for the GUI / mesh packages the reference imports at module level but the query path never calls.

    python oracle/build_ref.py          # build + stage (needs /root/reference; ~3-5 min, nvcc cross-compiles sm_100)
    oracle.build_ref.load()             # import the built extension (GPU box or here)
    oracle.build_ref.load_python()      # -> (modules, diff_operators, utils) of the staged reference
"""
from __future__ import annotations

import importlib.util
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/contour_cuda"
OUT = os.path.join(HERE, "_ref")
NAME = "ref_contour_cuda"


def so_path():
    return os.path.join(OUT, "build", NAME + ".so")


def build(verbose=False):
    if not os.path.isdir(REF_SRC):
        raise FileNotFoundError(REF_SRC + " (the reference is only mounted in the build container)")
    src = os.path.join(OUT, "src")
    bld = os.path.join(OUT, "build")
    os.makedirs(os.path.join(src, "includes"), exist_ok=True)
    os.makedirs(os.path.join(src, "glm"), exist_ok=True)
    os.makedirs(bld, exist_ok=True)
    for f in ["ext.cpp", "isocontours.cu", "intersection.cu", "includes/utils.h"]:
        text = open(os.path.join(REF_SRC, f)).read()
        text = re.sub(r"AT_DISPATCH_FLOATING_TYPES\((\w+)\.type\(\)", r"AT_DISPATCH_FLOATING_TYPES(\1.scalar_type()", text)
        open(os.path.join(src, f), "w").write(text)
    open(os.path.join(src, "glm", "glm.hpp"), "w").write("// empty stand-in: the reference includes glm but never uses it\n")
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=[os.path.join(src, f) for f in ["ext.cpp", "isocontours.cu", "intersection.cu"]],
         extra_include_paths=[src, os.path.join(src, "includes")], build_directory=bld, verbose=verbose,
         extra_cuda_cflags=["-O3"])
    return so_path()


def load():
    """Import the built reference extension -> module with extract_isocontours / get_fields_intersection_inter_slice."""
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    p = so_path()
    if not os.path.exists(p):
        raise FileNotFoundError(p + " -- run `python oracle/build_ref.py` in the build container")
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


REF_ROOT = "/root/reference"
PY_FILES = ["utils.py", "modules.py", "diff_operators.py", "level_collision.py", "sdf_meshing.py", "toolpath_utils.py",
            "vis.py", "loss_functions.py"]
DATA_FILES = ["sdf_data/bunnyz.xyz", "sdf_data/fertilityz.xyz", "data/rect_nozzle_shell.xyz"]
# imported at module level by the staged files, absent from this image, never called on the query path
STAND_INS = ["h5py", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "pyvista", "pyvistaqt",
             "skimage", "skimage.measure", "plyfile", "trimesh", "potpourri3d", "configargparse", "open3d", "networkx"]


def py_dir():
    return os.path.join(OUT, "py")


def python_staged():
    return os.path.exists(os.path.join(py_dir(), "utils.py"))


def stage_python():
    """Byte-copy the reference's Python query path and its input point clouds into oracle/_ref/py (git-ignored)."""
    if not os.path.isdir(REF_ROOT):
        raise FileNotFoundError(REF_ROOT + " (the reference is only mounted in the build container)")
    dst = py_dir()
    os.makedirs(dst, exist_ok=True)
    for f in PY_FILES:
        if os.path.exists(os.path.join(REF_ROOT, f)):
            shutil.copyfile(os.path.join(REF_ROOT, f), os.path.join(dst, f))
    tm = os.path.join(dst, "torchmeta")
    if os.path.isdir(tm):
        shutil.rmtree(tm)
    shutil.copytree(os.path.join(REF_ROOT, "torchmeta"), tm, ignore=shutil.ignore_patterns("__pycache__", "tests"))
    for f in DATA_FILES:
        os.makedirs(os.path.dirname(os.path.join(dst, f)), exist_ok=True)
        shutil.copyfile(os.path.join(REF_ROOT, f), os.path.join(dst, f))
    return dst


def load_python(cpu_shim=False):
    """Import the staged, unmodified reference modules -> (modules, diff_operators, utils).

    cpu_shim=True makes Tensor.cuda() / Module.cuda() the identity (the reference hard-codes .cuda() in its chunk
    loops, utils.py:161,175,236,261): that is how its own path runs on the host cores for the CPU baseline."""
    import types
    import torch
    if not python_staged():
        raise FileNotFoundError(py_dir() + " -- run `python oracle/build_ref.py` in the build container")
    for name in STAND_INS + ["qupkg", "qupkg.extract_isocontours"]:
        if name in sys.modules:
            continue
        try:
            if not name.startswith("qupkg"):
                importlib.import_module(name)
                continue
        except Exception:
            pass
        m = types.ModuleType(name)
        m.__getattr__ = lambda attr, _n=name: (_ for _ in ()).throw(AttributeError(attr)) if attr.startswith("__") else object
        sys.modules[name] = m
    if py_dir() not in sys.path:
        sys.path.insert(0, py_dir())
    if cpu_shim:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.empty_cache = lambda: None
    mods = []
    for name in ("modules", "diff_operators", "utils"):
        cached = sys.modules.get(name)
        if cached is not None and not str(getattr(cached, "__file__", "")).startswith(py_dir()):
            del sys.modules[name]      # a drop-in of the same name was imported earlier: the oracle wants the reference's
        mods.append(importlib.import_module(name))
    return tuple(mods)


if __name__ == "__main__":
    if not os.path.exists(so_path()) or "--force" in sys.argv:
        print(build(verbose="-v" in sys.argv))
    print(stage_python())
