"""Record goldens from the REFERENCE's compiled contour_cuda extension (oracle/_ref, built by oracle/build_ref.py) on a
B200 -- the only place it can run.  Run on the GPU box from the repo root:

    gpurun -- 'python tests/golden/make_contour_golden.py'        # writes gpurun_out/contour_ref_*.npz

then copy gpurun_out/contour_ref_*.npz into tests/golden/ and commit them.  tests/test_gpu_contours.py compares the CUDA
path with these files, so the comparison against the real reference kernels survives on a box where oracle/_ref (git-
ignored) is absent.  The reference is non-deterministic in slot order (isocontours.cu:85) and in the winner of a cell
(intersection.cu:119-124): the goldens hold ONE valid outcome and the tests compare modulo exactly those freedoms.

Recorded:
  contour_ref_isocontours.npz   torus(160,80) fixture, K = 40: used rows of contours_merged per isovalue (row stream up to
                                the last separator), contour_nums, and the derived segment counts S_k
  contour_ref_intersection.npz  (a) D = 24 fixture with 20 % NaN voxels: mask, indices, points
                                (b) the reference's own size (slice_toolpath.py:125: D = 255, Ns = 500, N1 = N2 = 128):
                                    packed mask bits, SHA-256 of mask and index tensors, points of every 997th cell
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "inf-3dp_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def field_of(V):
    return (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)


def reference_size_inputs():
    """The reference's own a6 size (slice_toolpath.py:125): D = 255, Ns = 500, N1 = N2 = 128; sin-free fields so that the
    inputs are bit-reproducible on any host (inf3dp.testing.corner_expanded_fields(poly=True))."""
    from inf3dp.testing import corner_expanded_fields
    fields, centers = corner_expanded_fields(255, nan_frac=0.0, poly=True)
    isos = [np.linspace(-0.9, 0.9, 500).astype(np.float32), np.linspace(-0.8, 0.8, 128).astype(np.float32),
            np.linspace(-0.8, 0.8, 128).astype(np.float32)]
    return fields, isos, centers


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.cpu().numpy()).tobytes()).hexdigest()


def main():
    from oracle import build_ref
    from inf3dp.testing import torus_mesh, rows_used, corner_expanded_fields
    ref = build_ref.load()
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")

    V, F = torus_mesh(160, 80)
    f = field_of(V)
    iso = np.linspace(f.min(), f.max(), 40).astype(np.float32)
    rm, rn = ref.extract_isocontours(*(torch.from_numpy(a).to(dev) for a in (V, F, f, iso)))
    rm, rn = rm.cpu().numpy(), rn.cpu().numpy()
    used = np.array([rows_used(rm[k], rn[k]) for k in range(len(iso))], np.int64)
    rows = np.concatenate([rm[k, :used[k]] for k in range(len(iso))], 0)
    assert all(not rm[k, used[k]:].any() for k in range(len(iso)))     # zero padding behind the used rows
    np.savez_compressed(os.path.join(out_dir, "contour_ref_isocontours.npz"), nu=160, nv=80, V=V, fields=f, iso=iso, rows=rows,
                        rows_used=used, contour_nums=rn, seg_counts=(used - rn).astype(np.int32),
                        device=torch.cuda.get_device_name(0), torch=torch.__version__)
    print("isocontours:", int(used.sum()), "rows,", int(rn.sum()), "loops")

    (s, a, b), centers = corner_expanded_fields(24, poly=True)
    si = np.linspace(-0.9, 0.9, 40).astype(np.float32)
    ai = np.linspace(-0.8, 0.8, 12).astype(np.float32)
    bi = np.linspace(-0.8, 0.8, 12).astype(np.float32)
    m, i, p = ref.get_fields_intersection_inter_slice(*(torch.from_numpy(x).to(dev) for x in (s, a, b, si, ai, bi, centers)))
    small = {"small_mask": m.cpu().numpy(), "small_idx": i.cpu().numpy(), "small_pts": p.cpu().numpy(),
             "small_si": si, "small_ai": ai, "small_bi": bi}

    fields, isos, centers = reference_size_inputs()
    m, i, p = ref.get_fields_intersection_inter_slice(*(torch.from_numpy(x).to(dev) for x in (*fields, *isos, centers)))
    torch.cuda.synchronize()
    cells = torch.arange(0, m.numel(), 997, device=dev)
    np.savez_compressed(os.path.join(out_dir, "contour_ref_intersection.npz"), **small,
                        big_mask_bits=np.packbits(m.cpu().numpy().ravel()), big_mask_sha=sha(m), big_idx_sha=sha(i),
                        big_cells=cells.cpu().numpy(), big_pts=p.view(-1, 3)[cells].cpu().numpy(),
                        big_shape=np.array([255, 500, 128, 128]), device=torch.cuda.get_device_name(0))
    print("intersection: small", int(small["small_mask"].sum()), "cells, reference size", int(m.sum()), "cells")


if __name__ == "__main__":
    main()
