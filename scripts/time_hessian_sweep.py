"""Scratch: the bench's 16M-point Hessian / curvature sweep alone (bench.bench_hessian_sweep)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import bench, inf3dp
dev = torch.device("cuda:0")
net = inf3dp.SingleBVPNet(type="sine", in_features=3)
net.load_state_dict({f"net.net.{i}.0.{k}": torch.from_numpy(np.asarray(v)) for i, (W, b) in enumerate(bench.golden_layers())
                     for k, v in (("weight", W), ("bias", b))})
net.to(dev)
print(json.dumps(bench.bench_hessian_sweep(net, dev, bench.load_peaks(), 0, 1)))
