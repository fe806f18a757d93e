"""Generate tests/golden/volume_golden.npz by running the REFERENCE's own grid helpers (torch CPU) in the build
container (/root/reference is mounted here; the GPU box has not got it -- the committed .npz is what travels):

  * sdf_meshing.gen_voxle_queries (sdf_meshing.py:46-65): rows of the N=37 grid, SHA-256 of the whole N=128 tensor
    (above 2^24 = 256^3 the float32 index rounding of the true division starts to matter: N=300 rows are sampled too)
  * toolpath_utils.extract_corner_fields (toolpath_utils.py:986-1019) on a random 7x7x7 grid
  * toolpath_utils.get_voxel_centers (toolpath_utils.py:1021-1033) for N=9 whole, SHA-256 for N=256

    python tests/golden/make_golden_volume.py
"""
import contextlib
import hashlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402
from make_golden_flows import StubFinder  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.empty_cache = lambda: None
    import_reference()
    for name in [n for n in sys.modules if n.split(".")[0] in StubFinder.TOPS]:
        del sys.modules[name]
    sys.meta_path.insert(0, StubFinder())
    with contextlib.redirect_stdout(io.StringIO()):
        import sdf_meshing
        import toolpath_utils
    out = {}
    q37, _, vs37 = sdf_meshing.gen_voxle_queries(37)
    out["queries37"] = q37.numpy()
    out["queries37_voxel_size"] = np.float64(vs37)
    q128, _, _ = sdf_meshing.gen_voxle_queries(128)
    out["queries128_sha256"] = np.array(sha(q128.numpy()))
    q300, _, _ = sdf_meshing.gen_voxle_queries(300)
    rows = np.concatenate([np.arange(0, 4096), np.arange(300 ** 3 - 4096, 300 ** 3),
                           np.random.default_rng(5).integers(0, 300 ** 3, 8192)])
    out["queries300_rows"] = rows.astype(np.int64)
    out["queries300"] = q300.numpy()[rows]
    out["queries300_sha256"] = np.array(sha(q300.numpy()))
    q_half, _, _ = sdf_meshing.gen_voxle_queries(21, cube_size=0.5)
    out["queries21_half"] = q_half.numpy()
    rng = np.random.default_rng(9)
    grid = torch.from_numpy(rng.standard_normal((7, 7, 7)).astype(np.float32))
    out["corner_grid"] = grid.numpy()
    out["corner_fields"] = toolpath_utils.extract_corner_fields(grid).numpy()
    out["centers9"] = toolpath_utils.get_voxel_centers(9).numpy()
    c256 = toolpath_utils.get_voxel_centers(256).numpy()
    out["centers256_sha256"] = np.array(sha(c256))
    out["centers256_diag"] = c256[np.arange(255), np.arange(255), np.arange(255)]
    np.savez_compressed(os.path.join(HERE, "volume_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "volume_golden.npz"), {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
