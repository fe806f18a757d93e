"""Scratch: reverse-mode engine timing on 4 M random points (order 1) and its parity against the FP32 SIMT engine."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp
torch.manual_seed(2048)
net = inf3dp.SingleBVPNet(type="sine", in_features=3).cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
x = (torch.rand(n, 3, device="cuda") * 2 - 1)
for _ in range(2): net.query(x, order=1, mode="tc_reverse")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): y, J, _ = net.query(x, order=1, mode="tc_reverse")
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"tc_reverse n={n}: {ms:.3f} ms  {n/ms/1e3:.1f} Mpts/s", flush=True)
m = 1 << 18
ys, Js, _ = net.query(x[:m], order=1, mode="simt")
cosang = torch.nn.functional.cosine_similarity(J[:m, 0].double(), Js[:, 0].double(), dim=1).clamp(-1, 1)
cr = torch.linalg.norm(torch.cross(J[:m, 0].double(), Js[:, 0].double(), dim=1), dim=1)
ang = torch.rad2deg(torch.atan2(cr, (J[:m, 0].double() * Js[:, 0].double()).sum(1)))
print(f"vs simt on {m} points: |dy| max {float((y[:m] - ys).abs().max()):.2e}  angle max {float(ang.max()):.4f} deg", flush=True)
