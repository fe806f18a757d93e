"""Design probe (CPU, torch): Hessian by FORWARD-over-REVERSE (3 tangent rows through the forward + backward sweep of the
reverse-mode engine) with FP16 hi/lo split tensor-core operands, against fp64 autograd.  Operand rounding only
(products / accumulation in fp32), the scales of the engine (power-of-two weight scale S, range scale G, tangent scale kT).

usage: python scripts/emu_hess_for.py [n_points]
"""
import math, sys
import torch

H = 256


def init_net(seed):
    torch.manual_seed(seed)
    Ws, bs = [], []
    dims = [3, H, H, H, H, 1]
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


def mm(a, Wt, split=True):
    if not split:
        return a @ Wt
    ah, al = split16(a); wh, wl = split16(Wt)
    return ah @ wh + al @ wh + ah @ wl


def pow2_scale(m, top=16.0):
    e = math.frexp(m)[1]
    return 2.0 ** (math.log2(top) - e)


def ref64(Ws, bs, x):
    x = x.double().requires_grad_(True)
    a = x
    for l in range(4):
        a = torch.sin(30 * (a @ Ws[l].double().T + bs[l].double()))
    y = a @ Ws[4].double().T + bs[4].double()
    g, = torch.autograd.grad(y.sum(), x, create_graph=True)
    Hs = [torch.autograd.grad(g[:, i].sum(), x, retain_graph=True)[0] for i in range(3)]
    return y.detach(), g.detach(), torch.stack(Hs, 1).detach()


def for_emul(Ws, bs, x, split=True, kT=2.0 ** -5):
    N = x.shape[0]
    W0 = 30 * Ws[0]; b0 = 30 * bs[0]
    z = x @ W0.T + b0
    a = torch.sin(z); c = torch.cos(z)
    zd = W0.T[None, :, :].expand(N, 3, H)      # tangents of the layer-0 arguments
    ad = c[:, None, :] * zd
    cs = [c]; cds = [-a[:, None, :] * zd]
    for l in (1, 2, 3):
        Wp = 30 * Ws[l]; S = pow2_scale(Wp.abs().max().item()); Wt = (Wp * S).T.contiguous()
        z = mm(a, Wt, split) / S + 30 * bs[l]
        zd = mm(ad.reshape(-1, H), Wt, split).reshape(N, 3, H) / S
        a = torch.sin(z); c = torch.cos(z)
        ad = c[:, None, :] * zd
        cs.append(c); cds.append(-a[:, None, :] * zd)
    y = a @ Ws[4].T + bs[4]
    G = pow2_scale(Ws[4].abs().max().item(), top=2.0)
    w4 = Ws[4][0] * G
    g = w4[None, :] * cs[3]                    # G g_3
    gd = kT * w4[None, None, :] * cds[3]       # kT G dg_3/dx_i
    mg = 0.0
    for l in (3, 2, 1):
        Wp = 30 * Ws[l]; S = pow2_scale(Wp.abs().max().item()); Wm = (Wp * S).contiguous()
        mg = max(mg, gd.abs().max().item())
        ga = mm(g, Wm, split) / S
        gda = mm(gd.reshape(-1, H), Wm, split).reshape(N, 3, H) / S
        g = ga * cs[l - 1]
        gd = gda * cs[l - 1][:, None, :] + kT * ga[:, None, :] * cds[l - 1]
    grad = (g @ W0) / G
    Hm = (gd.reshape(-1, H) @ W0).reshape(N, 3, 3) / (G * kT)
    return y, grad, Hm, mg


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    for seed in (2048, 7):
        Ws, bs = init_net(seed)
        x = torch.rand(N, 3) * 2 - 1
        y64, g64, H64 = ref64(Ws, bs, x)
        for name, split in (("fp32", False), ("fp16x3", True)):
            y, g, Hm, mg = for_emul(Ws, bs, x, split)
            relF = (Hm.double() - H64).flatten(1).norm(dim=1) / H64.flatten(1).norm(dim=1)
            asym = (Hm - Hm.transpose(1, 2)).abs().max().item()
            print(f"seed {seed} {name:7s}: H rel-Frobenius med={relF.median():.2e} p99.9={relF.quantile(0.999):.2e} max={relF.max():.2e}  "
                  f"max |kT G dg|={mg:.2f}  asymmetry={asym:.2e}  |H| med={H64.flatten(1).norm(dim=1).median():.1f}")
