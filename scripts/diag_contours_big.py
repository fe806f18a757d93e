"""Scratch: per-isovalue comparison of the CUDA contour path with the oracle on the F = 4 M torus."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp
from oracle import contour_oracle as co
nu, nv = int(sys.argv[1]), int(sys.argv[2])
drop = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
V, F = co.torus_mesh(nu, nv)
f = (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)
if drop > 0:
    rng = np.random.default_rng(17)
    F = np.ascontiguousarray(F[rng.random(len(F)) >= drop])
lo, hi = float(f.min()), float(f.max())
iso = np.linspace(lo + 0.08 * (hi - lo), hi - 0.08 * (hi - lo), 10).astype(np.float32)
mo, no, so_ = co.extract_isocontours(V, F, f, iso)
t = [torch.from_numpy(a).cuda() for a in (V, F, f, iso)]
m, n, s = inf3dp.extract_isocontours_raw(*t, want_seg_counts=True)
m, n, s = m.cpu().numpy(), n.cpu().numpy(), s.numpy()
for k in range(len(iso)):
    used_o = co.rows_used(mo[k], no[k]); used = co.rows_used(m[k], n[k])
    eq = np.array_equal(m[k], mo[k])
    msg = f"k={k} S={s[k]} (oracle {so_[k]}) loops={n[k]} (oracle {no[k]}) rows {used}/{used_o} equal={eq}"
    if not eq:
        diff = np.nonzero((m[k] != mo[k]).any(1))[0]
        a = m[k, :used]; b = mo[k, :used_o]
        sa = np.sort(a.view([('x', 'f4'), ('y', 'f4'), ('z', 'f4')]).ravel()); sb = np.sort(b.view([('x', 'f4'), ('y', 'f4'), ('z', 'f4')]).ravel())
        msg += f" differing rows={len(diff)} first={diff[:4]} last={diff[-2:]} same multiset={np.array_equal(sa, sb)}"
        lo_ = [len(l) for l in co.split_rows(m[k], n[k])]; lo2 = [len(l) for l in co.split_rows(mo[k], no[k])]
        msg += f"\n    ours loop lengths {lo_}\n    oracle loop lengths {lo2}"
    print(msg, flush=True)
