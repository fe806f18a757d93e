"""Scratch: marching cubes on the bench's volume (random-init SIREN at 512^3, median level) for ncu / timing."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import bench, inf3dp
dev = torch.device("cuda:0")
N = int(os.environ.get("MC_N", "512"))
net = inf3dp.SingleBVPNet(type="sine", in_features=3)
net.load_state_dict({f"net.net.{i}.0.{k}": torch.from_numpy(np.asarray(v)) for i, (W, b) in enumerate(bench.golden_layers())
                     for k, v in (("weight", W), ("bias", b))})
net.to(dev)
sdf = inf3dp.grid_field(net, N, chunk=1 << 23)
level = float(sdf.flatten()[::4099].median())
h = 2.0 / (N - 1)
reps = int(os.environ.get("MC_REPS", "3"))
for _ in range(2):
    v, f, nrm, val = inf3dp.marching_cubes(sdf, level, (h, h, h))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    v, f, nrm, val = inf3dp.marching_cubes(sdf, level, (h, h, h))
e1.record(); torch.cuda.synchronize()
print(f"marching cubes N={N} level={level:.5f}: {e0.elapsed_time(e1) / reps:.3f} ms verts={len(v)} faces={len(f)}")
print("checksum", float(v.double().sum()), int(f.long().sum()), float(nrm.double().sum()), float(val.double().sum()))
