"""Scratch: the bench's headline launch alone -- reverse engine on the 512^3 grid (rows output), 3 timed launches."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import bench, inf3dp
from inf3dp.query import gen_voxel_queries
dev = torch.device("cuda:0")
net = inf3dp.SingleBVPNet(type="sine", in_features=3)
net.load_state_dict({f"net.net.{i}.0.{k}": torch.from_numpy(np.asarray(v)) for i, (W, b) in enumerate(bench.golden_layers())
                     for k, v in (("weight", W), ("bias", b))})
net.to(dev)
coords, _, _ = gen_voxel_queries(512, device=dev)
n = coords.shape[0]
for _ in range(2):
    net.query(coords, order=1, mode="tc_reverse")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    net.query(coords, order=1, mode="tc_reverse")
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"tc_reverse 512^3 grid: {ms:.2f} ms  {n/ms/1e3:.1f} Mpts/s", flush=True)
