"""Small driver for ncu launch lists of the contour / intersection ops."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp
from oracle import contour_oracle as co
what = sys.argv[1] if len(sys.argv) > 1 else "iso"
if what == "iso":
    V, F = co.torus_mesh(1000, 500)
    f = (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)
    iso = np.linspace(f.min(), f.max(), 200).astype(np.float32)
    t = [torch.from_numpy(a).cuda() for a in (V, F, f, iso)]
    for _ in range(3):
        m, n = inf3dp.extract_isocontours_raw(*t)
    torch.cuda.synchronize()
    print("iso ok", int(n.sum()))
else:
    D, Ns, N1, N2 = 255, 500, 128, 128
    N = D + 1
    g = torch.stack(torch.meshgrid(*[torch.linspace(-1, 1, N, device="cuda")] * 3, indexing="ij"), -1)
    base = [g[..., 2] + 0.1 * torch.sin(3 * g[..., 0]), g[..., 0] + 0.2 * g[..., 1] ** 2 + 0.05 * torch.sin(5 * g[..., 2]),
            g[..., 1] - 0.15 * g[..., 2] * g[..., 0] + 0.05 * torch.sin(4 * g[..., 0])]
    offs = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
    sdf = (g.norm(dim=-1) - 0.8)
    fields = []
    for b in base:
        c = torch.stack([b[o[0]:o[0] + D, o[1]:o[1] + D, o[2]:o[2] + D] for o in offs], -1).contiguous()
        inside = torch.stack([sdf[o[0]:o[0] + D, o[1]:o[1] + D, o[2]:o[2] + D] for o in offs], -1).min(-1).values < 0
        c[~inside] = float("nan")
        fields.append(c)
    idx = torch.stack(torch.meshgrid(*[torch.arange(D, device="cuda", dtype=torch.float32)] * 3, indexing="ij"), -1)
    centers = ((idx + 0.5) * (2.0 / D) - 1.0).contiguous()
    si = torch.linspace(-0.9, 0.9, Ns, device="cuda")
    ai = torch.linspace(-0.8, 0.8, N1, device="cuda")
    bi = torch.linspace(-0.8, 0.8, N2, device="cuda")
    import time
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m, i, p = inf3dp.get_fields_intersection_inter_slice(*fields, si, ai, bi, centers)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    nb = 108 * D ** 3 + 4 * (Ns + N1 + N2) + 19 * Ns * N1 * N2
    print(f"isect ok hits={int(m.sum())} {dt*1e3:.3f} ms  {nb/dt/1e9:.0f} GB/s algorithmic")
