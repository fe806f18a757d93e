"""Scratch: timings of the volume row (marching cubes at N^3, intersection on grids vs expanded inputs)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for N in (256, 512):
    ax = torch.linspace(-1, 1, N, device="cuda")
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    vol = (torch.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.7 + 0.05 * torch.sin(9 * X) * torch.cos(7 * Y)).contiguous()
    del X, Y, Z
    v, f, n, val = inf3dp.marching_cubes(vol, 0.0, (2 / (N - 1),) * 3)
    ms = timeit(lambda: inf3dp.marching_cubes(vol, 0.0, (2 / (N - 1),) * 3))
    nb = 4 * N ** 3
    print(f"marching cubes N={N}: {ms:.3f} ms  verts={len(v)} faces={len(f)}  volume bytes {nb/1e6:.0f} MB -> {nb/ms/1e6:.0f} GB/s (one pass of the grid)")
    del vol

N = 256
g = torch.stack(torch.meshgrid(*[torch.linspace(-1, 1, N, device="cuda")] * 3, indexing="ij"), -1)
grids = [(g[..., 2] + 0.1 * torch.sin(3 * g[..., 0])).contiguous(), (g[..., 0] + 0.2 * g[..., 1] ** 2).contiguous(),
         (g[..., 1] - 0.15 * g[..., 2] * g[..., 0]).contiguous()]
Ns, N1, N2 = 500, 128, 128
si = torch.linspace(-0.9, 0.9, Ns, device="cuda"); ai = torch.linspace(-0.8, 0.8, N1, device="cuda"); bi = torch.linspace(-0.8, 0.8, N2, device="cuda")
ms_g = timeit(lambda: inf3dp.get_fields_intersection_from_grids(*grids, si, ai, bi, sync=False))
offs = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
D = N - 1
def expand(x):
    return torch.stack([x[o[0]:o[0] + D, o[1]:o[1] + D, o[2]:o[2] + D] for o in offs], -1).contiguous()
ms_x = timeit(lambda: [expand(x) for x in grids], reps=3)
exp = [expand(x) for x in grids]
idx = torch.stack(torch.meshgrid(*[torch.arange(D, device="cuda", dtype=torch.float32)] * 3, indexing="ij"), -1)
centers = ((idx + 0.5) * (2.0 / D) + (-1.0)).contiguous()
ms_e = timeit(lambda: inf3dp.get_fields_intersection_inter_slice(*exp, si, ai, bi, centers, sync=False))
a = inf3dp.get_fields_intersection_from_grids(*grids, si, ai, bi)
b = inf3dp.get_fields_intersection_inter_slice(*exp, si, ai, bi, centers)
same = torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(torch.nan_to_num(a[2]), torch.nan_to_num(b[2]))
cells = Ns * N1 * N2
print(f"intersection D={D}: grids {ms_g:.3f} ms | expanded {ms_e:.3f} ms (+ torch corner expansion {ms_x:.3f} ms) | identical={same} hits={int(a[0].sum())}/{cells}")
print(f"  contract bytes expanded {(108*D**3 + 19*cells)/1e9:.2f} GB -> {(108*D**3 + 19*cells)/ms_e/1e6:.0f} GB/s; grid-mode bytes {(12*N**3 + 19*cells)/1e9:.2f} GB -> {(12*N**3 + 19*cells)/ms_g/1e6:.0f} GB/s")
