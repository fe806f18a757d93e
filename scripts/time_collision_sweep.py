"""Scratch: the bench's 4M-waypoint collision sweep alone (bench.bench_collision), optionally with a torch profiler table."""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_194_304
print(json.dumps(bench.bench_collision(net, dev, 0, 1, n_waypoints=n)))
if "--profile" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        bench.bench_collision(net, dev, 0, 1, n_waypoints=1 << 20)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
