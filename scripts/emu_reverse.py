"""Design probe (CPU, torch): accuracy of a REVERSE-mode value+gradient SIREN evaluation with FP16 hi/lo split
tensor-core operands, against an fp64 evaluation.  Emulates operand rounding only (products/accumulation fp32).

usage: python scripts/emu_reverse.py [n_points]
"""
import math, sys
import numpy as np, torch

torch.manual_seed(2048)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
H = 256


def init_net(out=1):
    Ws, bs = [], []
    dims = [3, H, H, H, H, out]
    for i in range(5):
        lin = torch.nn.Linear(dims[i], dims[i + 1])
        Ws.append(lin.weight.detach().clone()); bs.append(lin.bias.detach().clone())
    for w in Ws:
        bound = math.sqrt(6 / w.shape[1]) / 30
        w.uniform_(-bound, bound)
    Ws[0].uniform_(-1 / 3, 1 / 3)
    return Ws, bs


def split16(t):
    hi = t.to(torch.float16).to(torch.float32)
    lo = (t - hi).to(torch.float16).to(torch.float32)
    return hi, lo


def mm(a, Wt, na, nw):
    """a [N,K] @ Wt [K,M]; na/nw = number of fp16 pieces kept for each operand (1 or 2); 0 = fp32 exact."""
    if na == 0 and nw == 0:
        return a @ Wt
    ah, al = split16(a) if na else (a, None)
    wh, wl = split16(Wt) if nw else (Wt, None)
    out = ah @ wh
    if na == 2:
        out = out + al @ wh
    if nw == 2:
        out = out + ah @ wl
    return out


def pow2_scale(m, top=16.0):
    e = math.frexp(m)[1]
    return 2.0 ** (math.log2(top) - e)


def reference64(Ws, bs, x):
    x = x.double().requires_grad_(True)
    a = x
    for l in range(4):
        a = torch.sin(30 * (a @ Ws[l].double().T + bs[l].double()))
    y = a @ Ws[4].double().T + bs[4].double()
    g, = torch.autograd.grad(y.sum(), x)
    return y.detach(), g.detach()


def reverse_emul(Ws, bs, x, fwd=(2, 2), bwd=(2, 2), stash16=False):
    W0 = 30 * Ws[0]; b0 = 30 * bs[0]
    z = x @ W0.T + b0
    a = torch.sin(z); cs = [torch.cos(z)]
    for l in (1, 2, 3):
        Wp = 30 * Ws[l]
        S = pow2_scale(Wp.abs().max().item())
        z = mm(a, (Wp * S).T.contiguous(), *fwd) / S + 30 * bs[l]
        a = torch.sin(z); c = torch.cos(z)
        if stash16:
            c = c.half().float()
        cs.append(c)
    y = a @ Ws[4].T + bs[4]
    G = pow2_scale(Ws[4].abs().max().item(), top=2.0)
    g = (Ws[4][0] * G)[None, :] * cs[3]          # G * dy/dz3
    for l in (3, 2, 1):
        Wp = 30 * Ws[l]
        S = pow2_scale(Wp.abs().max().item())
        ga = mm(g, (Wp * S).contiguous(), *bwd) / S   # g_a[l-1][k] = sum_n g_z[n] W'[n,k]
        g = ga * cs[l - 1]
    grad = (g @ W0) / G
    return y, grad, g.abs().max().item()


def forward_emul(Ws, bs, x, pieces=(2, 2)):
    W0 = 30 * Ws[0]; b0 = 30 * bs[0]
    z = x @ W0.T + b0
    a = torch.sin(z)
    da = torch.cos(z)[:, None, :] * W0.T[None, :, :]   # [N,3,H]
    for l in (1, 2, 3):
        Wp = 30 * Ws[l]
        S = pow2_scale(Wp.abs().max().item())
        Wt = (Wp * S).T.contiguous()
        z = mm(a, Wt, *pieces) / S + 30 * bs[l]
        dz = mm(da.reshape(-1, H), Wt, *pieces).reshape(-1, 3, H) / S
        a = torch.sin(z); da = torch.cos(z)[:, None, :] * dz
    y = a @ Ws[4].T + bs[4]
    J = da @ Ws[4][0]
    return y, J


def angle(u, v):
    u = u.double(); v = v.double()
    cr = torch.linalg.norm(torch.cross(u, v, dim=-1), dim=-1)
    return torch.rad2deg(torch.atan2(cr, (u * v).sum(-1)))


def report(name, y, g, y64, g64):
    ang = angle(g, g64)
    print(f"{name:46s} |df|max={float((y.double()-y64).abs().max()):.2e}  angle: med={float(ang.median()):.5f} "
          f"p99.9={float(ang.quantile(0.999)):.5f} max={float(ang.max()):.5f} deg", flush=True)


for seed in (2048, 7):
    torch.manual_seed(seed)
    Ws, bs = init_net()
    x = torch.rand(N, 3) * 2 - 1
    y64, g64 = reference64(Ws, bs, x)
    print(f"seed {seed}: N={N}  |grad| min={float(g64.norm(dim=1).min()):.3e} med={float(g64.norm(dim=1).median()):.3e}")
    y, g = forward_emul(Ws, bs, x, (0, 0)); report("fp32 forward-mode", y, g, y64, g64)
    y, g = forward_emul(Ws, bs, x, (2, 2)); report("forward-mode fp16x3", y, g, y64, g64)
    y, g = forward_emul(Ws, bs, x, (1, 1)); report("forward-mode fp16x1", y, g, y64, g64)
    y, g, gm = reverse_emul(Ws, bs, x, (0, 0), (0, 0)); report("reverse fp32", y, g, y64, g64)
    y, g, gm = reverse_emul(Ws, bs, x, (2, 2), (2, 2)); report(f"reverse fwd x3 / bwd x3 (max|Gg|={gm:.1f})", y, g, y64, g64)
    y, g, gm = reverse_emul(Ws, bs, x, (2, 2), (2, 1)); report("reverse fwd x3 / bwd g-split only (x2)", y, g, y64, g64)
    y, g, gm = reverse_emul(Ws, bs, x, (2, 2), (1, 2)); report("reverse fwd x3 / bwd W-split only (x2)", y, g, y64, g64)
    y, g, gm = reverse_emul(Ws, bs, x, (2, 2), (1, 1)); report("reverse fwd x3 / bwd x1", y, g, y64, g64)
    y, g, gm = reverse_emul(Ws, bs, x, (2, 2), (2, 2), stash16=True); report("reverse x3/x3, fp16 cos stash", y, g, y64, g64)
    y, g, gm = reverse_emul(Ws, bs, x, (2, 1), (2, 2)); report("reverse fwd a-split only / bwd x3", y, g, y64, g64)
    y, g, gm = reverse_emul(Ws, bs, x, (1, 1), (2, 2)); report("reverse fwd x1 / bwd x3", y, g, y64, g64)
