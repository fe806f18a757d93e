"""Scratch timing of the individual kernels (CUDA events); not the bench contract."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp
from oracle import contour_oracle as co

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

torch.manual_seed(2048)
net = inf3dp.SingleBVPNet(type="sine", in_features=3).cuda()
engines = [e for e in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["simt"]) if e != "none"]
n = 1 << 22
x = (torch.rand(n, 3, device="cuda") * 2 - 1)
for eng in engines:
    for order in (0, 1) + ((2,) if eng == "simt" else ()):
        nn = n if eng != "simt" else n // 4
        ms = timeit(lambda: net.query(x[:nn], order=order, mode=eng))
        print(f"siren[{eng}] order={order} n={nn}: {ms:.3f} ms  {nn/ms/1e3:.2f} Mpts/s", flush=True)

for nu, nv in ((500, 250), (1000, 500)):
    V, F = co.torus_mesh(nu, nv)
    f = (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)
    K = 200
    iso = np.linspace(f.min(), f.max(), K).astype(np.float32)
    t = [torch.from_numpy(a).cuda() for a in (V, F, f, iso)]
    ms = timeit(lambda: inf3dp.extract_isocontours_raw(*t, sync=False))
    nb = 12 * len(F) + 16 * len(V) + 4 * K + 12 * K * len(F) + 4 * K
    m, nums, segs = inf3dp.extract_isocontours_raw(*t, want_seg_counts=True)
    print(f"isocontours F={len(F)} K={K}: {ms:.3f} ms  {K/ms*1e3:.0f} layers/s  contract {nb/ms/1e6:.0f} GB/s  segs={int(segs.sum())} loops={int(nums.sum())}", flush=True)
