import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "inf-3dp_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "siren_golden.npz"))


def layers_from_golden(g, name):
    """[(W, b)] for a net stored whole in the golden file ('sdf', 'quat')."""
    import numpy as np
    out, i = [], 0
    while f"{name}/net.net.{i}.0.weight" in g:
        out.append((np.asarray(g[f"{name}/net.net.{i}.0.weight"]), np.asarray(g[f"{name}/net.net.{i}.0.bias"])))
        i += 1
    return out
