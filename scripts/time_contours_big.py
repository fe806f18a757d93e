"""Scratch: isocontour timing through the C ABI with preallocated buffers (as bench.py does), any mesh size.
usage: python scripts/time_contours_big.py nu nv K   (F = 2 * nu * nv faces; 2000 1000 -> SURVEY.md cfg 3 upper size F = 4 M)"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp
from oracle import contour_oracle as co
nu, nv, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
V, F = co.torus_mesh(nu, nv)
f = (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)
iso = np.linspace(f.min(), f.max(), K).astype(np.float32)
dev = torch.device("cuda:0")
t = [torch.from_numpy(a).to(dev) for a in (V, F, f, iso)]
nV, nF = len(V), len(F)
L = inf3dp._lib.lib()
P = inf3dp._lib.ptr
wsb = L.inf3dp_extract_isocontours_workspace_bytes(nV, nF, K)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
merged = torch.empty((K, nF, 3), dtype=torch.float32, device=dev)
nums = torch.empty(K, dtype=torch.int32, device=dev)
st = inf3dp._lib.stream_ptr(dev)
def call():
    inf3dp._lib.check(L.inf3dp_extract_isocontours(*[P(a) for a in t], nV, nF, K, P(merged), P(nums), P(ws), wsb, st))
for _ in range(3): call()
torch.cuda.synchronize()
reps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): call()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
seg = (ctypes.c_int32 * K)()
inf3dp._lib.check(L.inf3dp_extract_isocontours_status(P(ws), K, seg, st))
segs = np.frombuffer(seg, dtype=np.int32)
nb = 12 * nF + 16 * nV + 4 * K + 12 * K * nF + 4 * K
print(f"isocontours F={nF} K={K}: {ms:.3f} ms  {K/ms*1e3:.0f} layers/s  contract {nb/ms/1e6:.0f} GB/s ({nb/ms/1e6/6553.3*100:.1f} % of HBM peak)  "
      f"segs={int(segs.sum())} max S={int(segs.max())} loops={int(nums.sum())}", flush=True)
m = merged.cpu().numpy() if K * nF <= 2e8 else None
if m is not None:
    n = nums.cpu().numpy()
    for k in range(0, K, max(1, K // 7)):
        used = co.rows_used(m[k], int(n[k]))
        assert used == int(segs[k]) + int(n[k]), (k, used, int(segs[k]), int(n[k]))
    print("row accounting ok")
