"""Time the REFERENCE's compiled contour_cuda extension (oracle/_ref) on this GPU at the bench sizes -- the GPU comparator
BASELINE.md 5.6(b) calls "the number to beat".  Prints one JSON line; used by bench.py's `gpu_reference` leg too."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "inf-3dp_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def time_call(fn, reps):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    from oracle import build_ref
    from inf3dp.testing import torus_mesh, corner_expanded_fields
    ref = build_ref.load()
    dev = torch.device("cuda:0")
    out = {"device": torch.cuda.get_device_name(0)}
    V, F = torus_mesh(1000, 500)
    f = (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)
    for K in (200,):
        iso = np.linspace(f.min(), f.max(), K).astype(np.float32)
        t = [torch.from_numpy(a).to(dev) for a in (V, F, f, iso)]
        out[f"extract_isocontours_F1M_K{K}_ms"] = time_call(lambda: ref.extract_isocontours(*t), 1)
    (s, a, b), centers = corner_expanded_fields(255, nan_frac=0.0)
    isos = [np.linspace(-0.9, 0.9, 500).astype(np.float32), np.linspace(-0.8, 0.8, 128).astype(np.float32),
            np.linspace(-0.8, 0.8, 128).astype(np.float32)]
    t = [torch.from_numpy(x).to(dev) for x in (s, a, b, *isos, centers)]
    out["fields_intersection_D255_500x128x128_ms"] = time_call(lambda: ref.get_fields_intersection_inter_slice(*t), 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
