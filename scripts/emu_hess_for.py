"""Design probe (CPU, torch): Hessian by FORWARD-over-REVERSE (3 tangent rows through the forward + backward sweep of the
reverse-mode engine) with FP16 hi/lo split tensor-core operands, against fp64 autograd.  Operand rounding only.

usage: python scripts/emu_hess_for.py [n_points]
"""
import math, sys
import torch
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from emu_reverse import init_net, mm, pow2_scale  # noqa: E402  (runs its report at import: harmless, small N below)
