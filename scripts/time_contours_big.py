"""Scratch: isocontour timing at F = 4 M faces (SURVEY.md cfg 3 upper size)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp
from oracle import contour_oracle as co
nu, nv, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
V, F = co.torus_mesh(nu, nv)
f = (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)
iso = np.linspace(f.min(), f.max(), K).astype(np.float32)
t = [torch.from_numpy(a).cuda() for a in (V, F, f, iso)]
m, nums, segs = inf3dp.extract_isocontours_raw(*t, want_seg_counts=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): inf3dp.extract_isocontours_raw(*t, sync=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
nb = 12 * len(F) + 16 * len(V) + 4 * K + 12 * K * len(F) + 4 * K
print(f"isocontours F={len(F)} K={K}: {ms:.3f} ms  {K/ms*1e3:.0f} layers/s  contract {nb/ms/1e6:.0f} GB/s  segs={int(segs.sum())} max S={int(segs.max())} loops={int(nums.sum())}")
for k in range(0, K, max(1, K // 7)):
    used = co.rows_used(m[k].cpu().numpy(), int(nums[k]))
    assert used == int(segs[k]) + int(nums[k]), (k, used, int(segs[k]), int(nums[k]))
print("row accounting ok")
