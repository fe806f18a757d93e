"""Scratch timing of get_fields_intersection_inter_slice at the reference's size (D=255, Ns=500, N1=N2=128)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp
D, Ns, N1, N2 = 255, 500, 128, 128
N = D + 1
lin = torch.linspace(-1, 1, N, device="cuda")
gx, gy, gz = torch.meshgrid(lin, lin, lin, indexing="ij")
base = [gz + 0.1 * torch.sin(3 * gx), gx + 0.2 * gy ** 2, gy - 0.15 * gz * gx]
offs = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
fields = [torch.stack([b[o[0]:o[0] + D, o[1]:o[1] + D, o[2]:o[2] + D] for o in offs], -1).contiguous() for b in base]
idx = torch.stack(torch.meshgrid(*[torch.arange(D, device="cuda", dtype=torch.float32)] * 3, indexing="ij"), -1)
centers = ((idx + 0.5) * (2.0 / D) - 1.0).contiguous()
si = torch.linspace(-0.9, 0.9, Ns, device="cuda"); ai = torch.linspace(-0.8, 0.8, N1, device="cuda"); bi = torch.linspace(-0.8, 0.8, N2, device="cuda")
args = fields + [si, ai, bi, centers]
for _ in range(2):
    m, i, p = inf3dp.get_fields_intersection_inter_slice(*args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m, i, p = inf3dp.get_fields_intersection_inter_slice(*args, sync=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
B = 108 * D ** 3 + 4 * (Ns + N1 + N2) + 19 * Ns * N1 * N2
print(f"fields_intersection D={D} Ns={Ns} N1={N1} N2={N2}: {ms:.3f} ms, contract bytes {B/1e9:.3f} GB -> {B/ms/1e6:.0f} GB/s, cells hit {int(m.sum())}")
