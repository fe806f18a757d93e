"""Differential operators on fused SIREN outputs (reference: diff_operators.py).

`gradient`, `jacobian`, `divergence`, `laplace` are the reference's autograd formulations: they work unchanged
because `model_out` comes from a twice-differentiable custom Function whose backward is served by the fused
kernels (no graph through the layers is ever built).  `hessian` / `hessian3d` / `min_max_curvs_and_vectors`
recognise such outputs and fetch the Hessian from ONE order-2 launch instead of 1+3 reverse passes per channel
(diff_operators.py:6-44) and a per-point Python loop (:190-218).
"""
from __future__ import annotations

import torch
from torch.autograd import grad

from . import _lib


def _fused_source(y, x):
    """The module that produced `y` from `x` through inf3dp.siren._SirenValue, or None."""
    fn = getattr(y, "grad_fn", None)
    net = getattr(fn, "net", None)
    if net is None or not hasattr(net, "query"):
        return None
    try:
        src = fn.saved_tensors[0]
    except Exception:
        return None
    if src.data_ptr() != x.data_ptr() or src.shape != x.shape:
        return None
    return net


def gradient(y, x, grad_outputs=None):
    """diff_operators.py:128-132."""
    if grad_outputs is None:
        grad_outputs = torch.ones_like(y)
    return torch.autograd.grad(y, [x], grad_outputs=grad_outputs, create_graph=True)[0]


def hessian3d(y, x):
    """diff_operators.py:27-44: y [N,C], x [N,3] -> [N,C,3,3]."""
    net = _fused_source(y, x)
    if net is not None:
        return net.query(x, order=2)[2]
    n, ch = y.shape[0], y.shape[1]
    gy = torch.ones_like(y[..., 0])
    h = torch.zeros(n, ch, x.shape[-1], x.shape[-1], device=y.device)
    for i in range(ch):
        dydx = grad(y[..., i], x, gy, create_graph=True)[0]
        for j in range(x.shape[-1]):
            h[..., i, j, :] = grad(dydx[..., j], x, gy, create_graph=True)[0][..., :]
    return h


def hessian(y, x):
    """diff_operators.py:6-25: y [B,N,C], x [B,N,3] -> (h [B,N,C,3,3], status)."""
    net = _fused_source(y, x)
    if net is not None:
        h = net.query(x, order=2)[2]
    else:
        b, n = y.shape[:2]
        gy = torch.ones_like(y[..., 0])
        h = torch.zeros(b, n, y.shape[-1], x.shape[-1], x.shape[-1], device=y.device)
        for i in range(y.shape[-1]):
            dydx = grad(y[..., i], x, gy, create_graph=True)[0]
            for j in range(x.shape[-1]):
                h[..., i, j, :] = grad(dydx[..., j], x, gy, create_graph=True)[0][..., :]
    status = -1 if torch.any(torch.isnan(h)) else 0
    return h, status


def jacobian(y, x):
    """diff_operators.py:135-148: y [B,N,C], x [B,N,3] -> (jac [B,N,C,3], status)."""
    net = _fused_source(y, x)
    if net is not None:
        jac = net.query(x, order=1)[1]
    else:
        b, n = y.shape[:2]
        jac = torch.zeros(b, n, y.shape[-1], x.shape[-1], device=y.device)
        for i in range(y.shape[-1]):
            yf = y[..., i].view(-1, 1)
            jac[:, :, i, :] = grad(yf, x, torch.ones_like(yf), create_graph=True)[0]
    status = -1 if torch.any(torch.isnan(jac)) else 0
    return jac, status


def divergence(y, x):
    """diff_operators.py:121-125."""
    div = 0.0
    for i in range(y.shape[-1]):
        div = div + grad(y[..., i], x, torch.ones_like(y[..., i]), create_graph=True)[0][..., i:i + 1]
    return div


def laplace(y, x):
    """diff_operators.py:47-49."""
    return divergence(gradient(y, x), x)


def principal_curvatures(grad_y, hess):
    """Curvature tail on device: grad [N,3], hess [N,3,3] -> kappa [N,2] (min,max), dirs [N,2,3]
    (diff_operators.py:186-218 without the per-point loop; eigenvector signs are arbitrary, as with eigh)."""
    if not grad_y.is_cuda:
        raise RuntimeError("inf3dp: principal_curvatures needs CUDA tensors (no CPU fallback)")
    g = grad_y.detach().to(torch.float32).contiguous()
    h = hess.detach().to(torch.float32).contiguous()
    n = g.shape[0]
    kappa = torch.empty((n, 2), dtype=torch.float32, device=g.device)
    dirs = torch.empty((n, 2, 3), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib().inf3dp_principal_curvatures(_lib.ptr(g), _lib.ptr(h), n, _lib.ptr(kappa), _lib.ptr(dirs),
                                                          _lib.stream_ptr(g.device)))
    return kappa, dirs


def unit_normals(grads, eps=1e-12, out=None):
    """F.normalize(grads, dim=-1) for [..., 3] gradients on the device (toolpath_utils.py:180-181)."""
    if not grads.is_cuda:
        raise RuntimeError("inf3dp: unit_normals needs CUDA tensors (no CPU fallback)")
    g = grads.detach().to(torch.float32).contiguous()
    out = torch.empty_like(g) if out is None else out
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib().inf3dp_unit_normals(_lib.ptr(g), g.numel() // 3, float(eps), _lib.ptr(out),
                                                   _lib.stream_ptr(g.device)))
    return out


def unit_normals_rows(rows, eps=1e-12, out=None):
    """unit_normals on packed (value, d/dx, d/dy, d/dz) rows [n,4] -> [n,3]."""
    if not rows.is_cuda or rows.dtype != torch.float32 or not rows.is_contiguous() or rows.shape[-1] != 4:
        raise RuntimeError("inf3dp: unit_normals_rows needs a contiguous CUDA float32 [n,4] tensor (no CPU fallback)")
    n = rows.numel() // 4
    out = torch.empty((n, 3), dtype=torch.float32, device=rows.device) if out is None else out
    with torch.cuda.device(rows.device):
        _lib.check(_lib.lib().inf3dp_unit_normals_rows(_lib.ptr(rows), n, float(eps), _lib.ptr(out), _lib.stream_ptr(rows.device)))
    return out


def min_max_curvs_and_vectors(f, x_temp):
    """diff_operators.py:167-220: f maps [N,3] coords to the {'model_in','model_out'} dict of a fused network."""
    x_temp = x_temp.clone().detach().requires_grad_(True)
    out = f(x_temp)
    net = _fused_source(out["model_out"], out["model_in"])
    if net is None:
        raise RuntimeError("inf3dp: min_max_curvs_and_vectors needs a decoder built on inf3dp.SingleBVPNet")
    _, J, H = net.query(out["model_in"], order=2)
    return principal_curvatures(J[:, 0, :], H[:, 0])
