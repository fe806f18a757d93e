"""Scratch check of the reverse-mode engine against the fp64 oracle + timing."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp
from oracle import siren_oracle as so

torch.manual_seed(2048)
net = inf3dp.SingleBVPNet(type="sine", in_features=3).cuda()
sd = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
layers = so.params_from_state_dict(sd)
for n in (100, 128, 1000, 40000):
    x = torch.rand(n, 3, device="cuda") * 2 - 1
    y, J, _ = net.query(x, order=1, mode="tc_reverse")
    torch.cuda.synchronize()
    yo, Jo, _ = so.siren_jet(layers, x.cpu().numpy().astype(np.float64), order=1)
    dy = np.abs(y.cpu().numpy() - yo).max()
    ang = so.angle_deg(J.cpu().numpy()[:, 0], Jo[:, 0])
    rel = np.abs(np.linalg.norm(J.cpu().numpy()[:, 0], axis=1) / np.linalg.norm(Jo[:, 0], axis=1) - 1).max()
    print(f"n={n}: |df|max={dy:.3e} angle max={ang.max():.5f} deg  |grad| rel err max={rel:.2e}", flush=True)

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

n = 1 << 22
x = torch.rand(n, 3, device="cuda") * 2 - 1
for eng in sys.argv[1:] or ["tc_reverse"]:
    ms = timeit(lambda: net.query(x, order=1, mode=eng))
    print(f"siren[{eng}] order=1 n={n}: {ms:.3f} ms  {n/ms/1e3:.2f} Mpts/s", flush=True)
