"""Scratch: the three workloads whose kernels are profiled for profiles/r2_*: one bench-sized launch of the reverse engine
(512^3 grid), the intersection op at the reference's size, the 200-isovalue isocontour extraction (dense contract)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import bench, inf3dp
from inf3dp.query import gen_voxel_queries
from inf3dp.testing import torus_mesh, CORNER_OFFSETS
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "siren"):
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    net.load_state_dict({f"net.net.{i}.0.{k}": torch.from_numpy(np.asarray(v)) for i, (W, b) in enumerate(bench.golden_layers())
                         for k, v in (("weight", W), ("bias", b))})
    net.to(dev)
    coords, _, _ = gen_voxel_queries(512, device=dev)
    rows = torch.empty((coords.shape[0], 4), device=dev)
    for _ in range(2):
        net.query_rows(coords, out=rows)
    torch.cuda.synchronize()
    del coords, rows
if which in ("all", "isect"):
    D, Ns, N1, N2 = 255, 500, 128, 128
    lin = torch.linspace(-1, 1, D + 1, device=dev)
    gx, gy, gz = torch.meshgrid(lin, lin, lin, indexing="ij")
    base = [gz + 0.1 * torch.sin(3 * gx), gx + 0.2 * gy ** 2, gy - 0.15 * gz * gx]
    fields = [torch.stack([b[o[0]:o[0] + D, o[1]:o[1] + D, o[2]:o[2] + D] for o in CORNER_OFFSETS], -1).contiguous() for b in base]
    idx = torch.stack(torch.meshgrid(*[torch.arange(D, device=dev, dtype=torch.float32)] * 3, indexing="ij"), -1)
    centers = ((idx + 0.5) * (2.0 / D) - 1.0).contiguous()
    isos = [torch.linspace(-0.9, 0.9, Ns, device=dev), torch.linspace(-0.8, 0.8, N1, device=dev), torch.linspace(-0.8, 0.8, N2, device=dev)]
    for _ in range(2):
        inf3dp.get_fields_intersection_inter_slice(*fields, *isos, centers)
    del fields, centers, base
if which in ("all", "contours"):
    V, F = torus_mesh(1000, 500)
    f = (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)
    iso = np.linspace(f.min(), f.max(), 200).astype(np.float32)
    t = [torch.from_numpy(a).to(dev) for a in (V, F, f, iso)]
    for _ in range(2):
        inf3dp.extract_isocontours(*t)
torch.cuda.synchronize()
print("ok")
