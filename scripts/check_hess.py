"""Scratch check of the forward-over-reverse Hessian engine (mode tc_reverse, order 2) against the fp64 oracle and the
second-order forward jet (tc_precise), plus timing of both."""
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
for n in (5, 32, 33, 257, 1000, 20000):
    x = torch.rand(n, 3, device="cuda") * 2 - 1
    y, J, H = net.query(x, order=2, mode="tc_reverse")
    yj, Jj, Hj = net.query(x, order=2, mode="tc_precise")
    torch.cuda.synchronize()
    yo, Jo, Ho = so.siren_jet(layers, x.cpu().numpy().astype(np.float64), order=2)
    dy = np.abs(y.cpu().numpy() - yo).max()
    ang = so.angle_deg(J.cpu().numpy()[:, 0], Jo[:, 0]).max()
    Hn = H.cpu().numpy()[:, 0]; Hr = Ho[:, 0]
    rel = np.linalg.norm((Hn - Hr).reshape(n, -1), axis=1) / np.linalg.norm(Hr.reshape(n, -1), axis=1)
    relj = np.linalg.norm((Hj.cpu().numpy()[:, 0] - Hr).reshape(n, -1), axis=1) / np.linalg.norm(Hr.reshape(n, -1), axis=1)
    print(f"n={n}: |df|max={dy:.3e} angle max={ang:.5f} deg  H rel-Fro: for med={np.median(rel):.2e} max={rel.max():.2e} | jet max={relj.max():.2e}  sym={np.array_equal(Hn, np.swapaxes(Hn,-1,-2))}", flush=True)

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
for eng in ("tc_reverse", "tc_precise"):
    ms = timeit(lambda: net.query(x, order=2, mode=eng), iters=3, warm=1)
    print(f"siren[{eng}] order=2 n={n}: {ms:.3f} ms  {n/ms/1e3:.2f} Mpts/s", flush=True)
os.environ["INF3DP_TC_DEBUG"] = "1"
