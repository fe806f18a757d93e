"""Drop-in for the reference's `qupkg` package (contour_cuda/setup.py:8-12)."""
