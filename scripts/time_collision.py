"""Scratch: where the time of the collision sweep goes (per-engine rates on 16.7M points)."""
import os, sys, math, copy
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp
torch.manual_seed(2048)
nets = [inf3dp.SingleBVPNet(type="sine", in_features=3, out_features=o).cuda() for o in (1, 1, 1, 1, 4)]
level = inf3dp.MLPClassifier(num_classes=8).cuda()
def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
n = 1 << 24
x = torch.rand(n, 3, device="cuda") * 2 - 1
x4 = torch.rand(n, 4, device="cuda") * 2 - 1
print(f"sdf value+grad   : {n/timeit(lambda: nets[0].query(x, order=1))/1e3:8.1f} Mpts/s")
print(f"slice value only : {n/timeit(lambda: nets[1].query(x, order=0))/1e3:8.1f} Mpts/s")
print(f"quat value (out4): {n/timeit(lambda: nets[4].query(x, order=0))/1e3:8.1f} Mpts/s")
print(f"level MLP logits : {n/timeit(lambda: level(x4))/1e3:8.1f} Mpts/s")
m = torch.rand(n, device="cuda") < 0.5
print(f"mask nonzero+gather [n,3]: {timeit(lambda: x[m.nonzero(as_tuple=True)[0]]):.3f} ms")
