import os, sys
import numpy as np, torch
ROOT="/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp
from oracle import contour_oracle as co
V, F = co.torus_mesh(1000, 500)
f = (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)
K = 200
iso = np.linspace(f.min(), f.max(), K).astype(np.float32)
t = [torch.from_numpy(a).cuda() for a in (V, F, f, iso)]
m, nums, segs = inf3dp.extract_isocontours_raw(*t, want_seg_counts=True)
nums = nums.cpu().numpy(); segs = segs.numpy()
order = np.argsort(-nums)[:8]
print("most pieces:", [(int(k), int(nums[k]), int(segs[k])) for k in order])
print("k=189:", nums[189], segs[189], " k=188:", nums[188], segs[188])
