"""Drop-in `modules` for the reference scripts (`import modules`, exp_scripts/toolpath.py:3-6).

Put inf-3dp_b200/dropin ahead of the reference root on sys.path, or call inf3dp.patch_reference().
Only the classes on the accelerated hot path are provided.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inf3dp.siren import SingleBVPNet, MLPClassifier  # noqa: E402,F401
