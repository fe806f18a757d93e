"""Drop-in for the reference's compiled module `qupkg.extract_isocontours` (contour_cuda/ext.cpp:49-52).

`from qupkg.extract_isocontours import extract_isocontours, get_fields_intersection_inter_slice`
(toolpath_utils.py:16) resolves here when inf-3dp_b200/ is on sys.path ahead of any reference build.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inf3dp.contours import extract_isocontours, get_fields_intersection_inter_slice  # noqa: E402,F401
