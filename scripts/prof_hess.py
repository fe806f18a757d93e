"""Scratch: one 1 M-point order-2 query on the forward-over-reverse engine (and one on the second-order jet) for ncu."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import bench, inf3dp
dev = torch.device("cuda:0")
net = inf3dp.SingleBVPNet(type="sine", in_features=3)
net.load_state_dict({f"net.net.{i}.0.{k}": torch.from_numpy(np.asarray(v)) for i, (W, b) in enumerate(bench.golden_layers())
                     for k, v in (("weight", W), ("bias", b))})
net.to(dev)
x = torch.rand(1 << 20, 3, device=dev) * 2 - 1
for _ in range(2):
    net.query(x, order=2, mode="tc_reverse")
net.query(x, order=2, mode="tc_precise")
torch.cuda.synchronize()
print("ok")
