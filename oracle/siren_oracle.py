"""ORACLE (test infrastructure, not product code) -- numpy restatement of the reference SIREN hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The shipped path (inf-3dp_b200/inf3dp) never does.

What is restated (all citations relative to /root/reference):
  * network forward            modules.py:16-25 (BatchLinear: x @ W^T + b), modules.py:32-34 (Sine: sin(30 z)),
                               modules.py:119-160 (SingleBVPNet: `num_hidden_layers` sine layers, outermost linear)
  * spatial gradient           diff_operators.py:128-132 (autograd.grad of sum(y * grad_outputs) wrt coords)
  * Hessian                    diff_operators.py:27-44  (hessian3d: grad of each gradient component)
  * curvature tail             diff_operators.py:167-220, 280-298 (tangent basis, 2x2 shape operator, eigh)
  * level classifier           exp_scripts/train_levels.py:35-51 (Linear-ReLU-Linear-ReLU-Linear)
  * query-grid generator       sdf_meshing.py:46-65 (gen_voxle_queries, including the float-division shear)
  * point-cloud normalisation  sdf_meshing.py:17-38 (PCD.__init__)

The derivative is carried analytically (forward-mode jet), which is mathematically what the reference's
reverse-mode autograd evaluates.  Pinning: the reference ships no golden vectors, so this
restatement is pinned against outputs of the reference itself, generated in the build container by
tests/golden/make_golden.py (which imports /root/reference) and committed as tests/golden/siren_golden.npz.
"""
from __future__ import annotations

import numpy as np

OMEGA0 = 30.0  # modules.py:34, hard-coded in the reference


def params_from_state_dict(sd, prefix="net.net.", dtype=np.float64):
    """[(W, b)] per layer from a reference state_dict (keys net.net.{i}.0.weight/bias)."""
    layers = []
    i = 0
    while f"{prefix}{i}.0.weight" in sd:
        W = np.asarray(sd[f"{prefix}{i}.0.weight"], dtype=dtype)
        b = np.asarray(sd[f"{prefix}{i}.0.bias"], dtype=dtype)
        layers.append((W, b))
        i += 1
    if not layers:
        raise KeyError("no net.net.{i}.0.weight keys in state_dict")
    return layers


def siren_jet(layers, x, order=1, dtype=np.float64, omega0=OMEGA0):
    """Value / Jacobian / Hessian of the sine MLP wrt the 3 input coordinates.

    layers: [(W_l [out_l, in_l], b_l [out_l])], the last one linear (modules.py:139-140 outermost_linear=True).
    x: [N, 3].  Returns y [N, out], J [N, out, 3] (order>=1 else None), H [N, out, 3, 3] (order>=2 else None).
    """
    x = np.asarray(x, dtype=dtype)
    n, d = x.shape
    a = x
    da = None  # [N, d, features]
    d2a = None  # [N, d, d, features]
    L = len(layers)
    for li, (W, b) in enumerate(layers):
        W = np.asarray(W, dtype=dtype)
        b = np.asarray(b, dtype=dtype)
        last = li == L - 1
        z = a @ W.T + b  # modules.py:23-24
        if order >= 1:
            if da is None:
                dz = np.broadcast_to(W.T[None, :, :], (n, d, W.shape[0])).astype(dtype)  # dz_i = W[:, i]
            else:
                dz = da @ W.T
        if order >= 2:
            if d2a is None:
                d2z = np.zeros((n, d, d, W.shape[0]), dtype=dtype)
            else:
                d2z = d2a @ W.T
        if last:
            y = z
            J = np.transpose(dz, (0, 2, 1)) if order >= 1 else None
            H = np.transpose(d2z, (0, 3, 1, 2)) if order >= 2 else None
            return y, J, H
        w = np.asarray(omega0, dtype=dtype)
        s = np.sin(w * z)  # modules.py:34
        if order >= 1:
            c = np.cos(w * z)
            wc = w * c
            if order >= 2:
                d2a = wc[:, None, None, :] * d2z - (w * w) * s[:, None, None, :] * dz[:, :, None, :] * dz[:, None, :, :]
            da = wc[:, None, :] * dz
        a = s
    raise AssertionError("unreachable")


def relu_mlp(layers, x, dtype=np.float64):
    """exp_scripts/train_levels.py:35-51: Linear-ReLU-...-Linear logits."""
    a = np.asarray(x, dtype=dtype)
    for li, (W, b) in enumerate(layers):
        a = a @ np.asarray(W, dtype=dtype).T + np.asarray(b, dtype=dtype)
        if li != len(layers) - 1:
            a = np.maximum(a, 0)
    return a


def level_with_confidence(logits):
    """utils.py:294-295: softmax over logits, then (max prob, argmax)."""
    z = logits - logits.max(axis=-1, keepdims=True)
    p = np.exp(z)
    p /= p.sum(axis=-1, keepdims=True)
    return p.argmax(axis=-1), p.max(axis=-1)


def principal_curvatures(grad, hess):
    """diff_operators.py:167-220 restated without the per-point Python loop.

    grad [N,3], hess [N,3,3] -> kappa [N,2] (min,max), dirs [N,2,3].
    Tangent basis (:190-201): a = unit vector on the axis of the smallest |n| component (first index on
    ties, torch.argmin), e1 = normalise(a - (a.n) n), e2 = n x e1.  Shape operator (:203-207)
    W = -[[e1 H e1, e1 H e2],[e1 H e2, e2 H e2]] / |grad|.  eigh (:208) ascending eigenvalues; eigenvectors are
    determined up to sign, so callers compare directions modulo sign.
    """
    g = np.asarray(grad, dtype=np.float64)
    Hm = np.asarray(hess, dtype=np.float64)
    gn = np.linalg.norm(g, axis=1, keepdims=True)
    nrm = g / gn
    idx = np.argmin(np.abs(nrm), axis=1)
    a = np.zeros_like(nrm)
    a[np.arange(len(a)), idx] = 1.0
    e1 = a - (a * nrm).sum(1, keepdims=True) * nrm
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.cross(nrm, e1)
    He1 = np.einsum("nij,nj->ni", Hm, e1)
    He2 = np.einsum("nij,nj->ni", Hm, e2)
    w11 = -(e1 * He1).sum(1) / gn[:, 0]
    w12 = -(e1 * He2).sum(1) / gn[:, 0]
    w22 = -(e2 * He2).sum(1) / gn[:, 0]
    Wm = np.stack([np.stack([w11, w12], -1), np.stack([w12, w22], -1)], -2)
    ev, evec = np.linalg.eigh(Wm)
    pmin = evec[:, 0, 0, None] * e1 + evec[:, 1, 0, None] * e2
    pmax = evec[:, 0, 1, None] * e1 + evec[:, 1, 1, None] * e2
    return ev, np.stack([pmin, pmax], 1)


def gen_voxel_queries(N, cube_size=1.0):
    """sdf_meshing.py:46-65, including the true-division-on-LongTensor quirk.

    The reference computes float32 `samples[:,1] = (idx / N) % N` (true division), i.e. fractional indices;
    everything is float32 arithmetic after the int64 -> float division, which numpy reproduces:
    torch's `LongTensor / int` yields float32, `% N` is a float32 fmod-with-floor, and the scale/offset
    is float32 * python-float (rounded to float32) + python-float.
    """
    idx = np.arange(N ** 3, dtype=np.int64)
    voxel_size = cube_size * 2.0 / (N - 1)
    s2 = (idx % N).astype(np.float32)
    q1 = (idx.astype(np.float32) / np.float32(N)).astype(np.float32)
    s1 = np.mod(q1, np.float32(N)).astype(np.float32)
    q0 = (q1 / np.float32(N)).astype(np.float32)
    s0 = np.mod(q0, np.float32(N)).astype(np.float32)
    vs = np.float32(voxel_size)
    out = np.empty((N ** 3, 3), dtype=np.float32)
    out[:, 0] = s0 * vs + np.float32(-cube_size)
    out[:, 1] = s1 * vs + np.float32(-cube_size)
    out[:, 2] = s2 * vs + np.float32(-cube_size)
    return out


def normalise_point_cloud(xyz):
    """sdf_meshing.py:23-38 (keep_aspect_ratio=True): centre, scale by global min/max to [-0.975, 0.975]."""
    c = np.array(xyz, dtype=np.float64)
    c -= c.mean(axis=0, keepdims=True)
    cmax, cmin = c.max(), c.min()
    c = (c - cmin) / (cmax - cmin)
    c -= 0.5
    c *= 1.95
    return c


def angle_deg(u, v):
    """Angle between row vectors in degrees (the north_star gradient tolerance is stated on this)."""
    u = np.asarray(u, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    cr = np.linalg.norm(np.cross(u, v), axis=-1)
    dt = (u * v).sum(-1)
    return np.degrees(np.arctan2(cr, dt))
