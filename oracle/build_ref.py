"""ORACLE build recipe (test infrastructure) -- compile the REFERENCE contour_cuda extension into oracle/_ref/.

The reference's native path is three source files (contour_cuda/{ext.cpp,isocontours.cu,intersection.cu});
they are compiled from where they lie under /root/reference.  Nothing from the reference is committed:
oracle/_ref/ is git-ignored (it still travels to the GPU box with the gpurun snapshot).

Two build-only adaptations are applied to a scratch copy under oracle/_ref/src (behaviour unchanged):
  * `X.type()` -> `X.scalar_type()` in the AT_DISPATCH_FLOATING_TYPES calls (isocontours.cu:242,
    intersection.cu:258): torch >= 2.x no longer converts DeprecatedTypeProperties to ScalarType
  * an empty glm/glm.hpp: includes/utils.h:2 includes it, nothing uses it, and the reference's
    third_party/glm is not vendored (setup.py:17)
The module is used by tests/ (parity on the GPU) and tests/golden/make_contour_golden.py only.

    python oracle/build_ref.py          # build (needs /root/reference; ~3-5 min, nvcc cross-compiles sm_100)
    oracle.build_ref.load()             # import the built module (GPU box or here)
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


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
