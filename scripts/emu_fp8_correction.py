"""Design probe (CPU, torch): can the two correction MMAs of the fp16x3 scheme (a_lo W_hi + a_hi W_lo) be replaced by ONE
FP8 MMA (kind::f8f6f4, K = 32 per instruction = [a_lo8 | a_hi8] x [W_hi8 ; W_lo8]) that accumulates into the same FP32
accumulator?  That would cut the tensor work per K-step from 3 to 2 MMA issues (ceiling 50 % instead of 33 %).

Result (100 k points, two seeds): maximum gradient angle error 0.09 - 0.24 degrees (median 0.002) -- the 0.1 degree parity
bound is a bound on the MAXIMUM, so the scheme is NOT parity grade and was not built.  fp16x3 on the same points: 0.003 - 0.005.

usage: python scripts/emu_fp8_correction.py [n_points]
"""
import math, sys, importlib.util
import torch
sys.argv = [sys.argv[0]] + sys.argv[1:]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
import os
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu_reverse.py')).read().split("for seed in (2048, 7):")[0]
exec(src)
E4, E5 = torch.float8_e4m3fn, torch.float8_e5m2
def q(t, dt, sc):
    return (t * sc).to(dt).to(torch.float32) / sc
MODE = {}
def mm8(a, Wt, a_dt=E4, w_dt=E5, a_lo_sc=2.0**20, a_hi_sc=2.0**8, wtop=2.0**15):
    # Wt is already scaled by S with max in [8,16); rescale so max ~ wtop for fp16 range realism (power of two: no effect on fp16 rounding)
    ah, al = split16(a)
    wh, wl = split16(Wt)
    m = Wt.abs().max().item()
    r = 2.0 ** (math.floor(math.log2(wtop / m)))
    # quantised correction operands
    al8 = q(al, a_dt, a_lo_sc)
    ah8 = q(ah, a_dt, a_hi_sc)
    wh8 = q(wh * r, w_dt, 1.0 / a_lo_sc) / r
    wl8 = q(wl * r, w_dt, 1.0 / a_hi_sc) / r
    return ah @ wh + (al8 @ wh8 + ah8 @ wl8)

def reverse8(Ws, bs, x, fwd8=True, bwd8=True, **kw):
    W0 = 30 * Ws[0]; b0 = 30 * bs[0]
    z = x @ W0.T + b0
    a = torch.sin(z); cs = [torch.cos(z)]
    for l in (1, 2, 3):
        Wp = 30 * Ws[l]
        S = pow2_scale(Wp.abs().max().item())
        Wt = (Wp * S).T.contiguous()
        z = (mm8(a, Wt, **kw) if fwd8 else mm(a, Wt, 2, 2)) / S + 30 * bs[l]
        a = torch.sin(z); cs.append(torch.cos(z))
    y = a @ Ws[4].T + bs[4]
    G = pow2_scale(Ws[4].abs().max().item(), top=2.0)
    g = (Ws[4][0] * G)[None, :] * cs[3]
    gm = 0
    for l in (3, 2, 1):
        Wp = 30 * Ws[l]
        S = pow2_scale(Wp.abs().max().item())
        Wm = (Wp * S).contiguous()
        if bwd8:
            # g has range: scale so max|g| row-wise? use global pow2 so |g| <= 1 like the kernel's G
            gs = 2.0 ** (-math.ceil(math.log2(g.abs().max().item())))
            ga = mm8(g * gs, Wm, **kw) / gs / S
        else:
            ga = mm(g, Wm, 2, 2) / S
        g = ga * cs[l - 1]
    grad = (g @ W0) / G
    return y, grad

for seed in (2048, 7):
    torch.manual_seed(seed)
    Ws, bs = init_net()
    x = torch.rand(N, 3) * 2 - 1
    y64, g64 = reference64(Ws, bs, x)
    y, g, gm = reverse_emul(Ws, bs, x, (2, 2), (2, 2)); report("reverse x3/x3", y, g, y64, g64)
    y, g = reverse8(Ws, bs, x); report("fp8corr fwd+bwd (A e4m3, W e5m2)", y, g, y64, g64)
    y, g = reverse8(Ws, bs, x, bwd8=False); report("fp8corr fwd only", y, g, y64, g64)
    y, g = reverse8(Ws, bs, x, fwd8=False); report("fp8corr bwd only", y, g, y64, g64)
    y, g = reverse8(Ws, bs, x, a_dt=E5, w_dt=E4, a_lo_sc=2.0**7, a_hi_sc=2.0**-5); report("fp8corr (A e5m2, W e4m3)", y, g, y64, g64)
