"""Drop-in `diff_operators` (reference diff_operators.py): same names, argument meaning and return shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inf3dp.diffops import (gradient, hessian, hessian3d, jacobian, divergence, laplace,  # noqa: E402,F401
                            min_max_curvs_and_vectors)
