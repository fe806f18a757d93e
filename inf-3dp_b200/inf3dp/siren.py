"""Drop-in SIREN field network backed by the fused sm_100a jet kernels.

Mirrors the reference surface (paths relative to the INF-3DP repository):
  * modules.py:119-160  SingleBVPNet(out_features, type, in_features, mode, hidden_features, num_hidden_layers)
                        forward({'coords': x}) -> {'model_in': x' (fresh leaf, requires_grad), 'model_out': y}
  * modules.py:622-635  sine_init / first_layer_sine_init (same RNG call order => same weights for a given seed)
  * state_dict keys     net.net.{i}.0.weight / .bias -- reference checkpoints load unchanged
  * diff_operators.py   gradient / hessian / hessian3d keep working on the outputs: `model_out` is produced by a
                        torch.autograd.Function whose backward contracts with the Jacobian computed by the same
                        fused kernel (forward-mode, no autograd graph through the layers) and whose
                        double-backward contracts with the Hessian from an order-2 launch.
  * exp_scripts/train_levels.py:35-51  MLPClassifier (ReLU level MLP), same keys network.{0,2,4}.*

Weights are constants for these kernels (inference); gradients flow to the coordinates only.
There is no CPU path: CPU tensors raise, as the reference's CHECK_CUDA does for its native ops.
"""
from __future__ import annotations

import ctypes
import math

import torch
from torch import nn

from . import _lib

_JET_MODE = {"mode": "auto"}


def set_engine(mode: str):
    """Select the execution engine for all SIREN queries: 'auto' | 'simt' | 'tc_precise' | 'tc_fast' | 'tc_reverse'."""
    if mode not in _lib.MODES:
        raise ValueError(f"unknown engine {mode!r}; expected one of {sorted(_lib.MODES)}")
    _JET_MODE["mode"] = mode


def get_engine() -> str:
    return _JET_MODE["mode"]


class PackedWeights:
    """Device blob produced by inf3dp_siren_pack_weights, re-packed when the source parameters change."""

    def __init__(self, in_features, hidden, num_hidden_layers, out_features, activation, omega0=30.0):
        self.cfg = (int(in_features), int(hidden), int(num_hidden_layers), int(out_features))
        self.activation = activation
        self.omega0 = float(omega0)
        self.blob = None
        self._sig = None

    def _release(self):
        """The library mirrors blob headers by device address: tell it before this blob's memory can be reused."""
        if self.blob is not None:
            try:
                _lib.lib().inf3dp_siren_release(_lib.ptr(self.blob))
            except Exception:
                pass
            self.blob = None

    def __del__(self):
        self._release()

    def __deepcopy__(self, memo):
        # a copied module re-packs on first use (its parameters are new tensors): never share or clone the blob
        return PackedWeights(*self.cfg, self.activation, self.omega0)

    def _signature(self, weights, biases):
        return tuple((t.data_ptr(), t._version, str(t.device)) for t in list(weights) + list(biases))

    def get(self, weights, biases):
        dev = weights[0].device
        if dev.type != "cuda":
            raise RuntimeError("inf3dp: the SIREN kernels need CUDA tensors (no CPU fallback); call .cuda() on the model")
        sig = self._signature(weights, biases)
        if self.blob is not None and sig == self._sig:
            return self.blob
        L = _lib.lib()
        nbytes = L.inf3dp_siren_packed_bytes(*self.cfg)
        if nbytes == 0:
            raise RuntimeError("inf3dp: unsupported network configuration in=%d hidden=%d hidden_layers=%d out=%d "
                               "(supported: hidden 256, in 3 or 4, out 1..64)" % self.cfg)
        ws = [w.detach().to(torch.float32).contiguous() for w in weights]
        bs = [b.detach().to(torch.float32).contiguous() for b in biases]
        n = len(ws)
        wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in ws])
        bp = (ctypes.c_void_p * n)(*[b.data_ptr() for b in bs])
        self._release()
        blob = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.inf3dp_siren_pack_weights(wp, bp, *self.cfg, self.activation, self.omega0, _lib.ptr(blob),
                                                   nbytes, _lib.stream_ptr(dev)))
        self.blob, self._sig = blob, sig
        return blob


_WORKSPACES = {}          # (device, stream) -> scratch tensor, least recently used first
_MAX_WORKSPACES = 8       # 38 MB each: a process that keeps creating streams must not accumulate them


def _jet_workspace(dev):
    """Scratch of the reverse-mode engine, one per (device, stream): launches on one stream are ordered, launches
    on different streams must not share it.  A small LRU: the oldest entry is dropped beyond _MAX_WORKSPACES (the
    caching allocator keeps its memory alive until the kernels queued on it have run)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(dev).cuda_stream)
    ws = _WORKSPACES.pop(key, None)
    if ws is None:
        with torch.cuda.device(dev):
            nbytes = _lib.lib().inf3dp_siren_jet_workspace_bytes()
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        while len(_WORKSPACES) >= _MAX_WORKSPACES:
            _WORKSPACES.pop(next(iter(_WORKSPACES)))      # allocated on its own stream: the allocator orders its reuse
    _WORKSPACES[key] = ws
    return ws


def jet(blob, x, order, out_features, mode=None):
    """Raw fused query: x [n, in] (CUDA, f32) -> y [n,out], J [n,out,3] | None, H [n,out,3,3] | None."""
    if not x.is_cuda:
        raise RuntimeError("inf3dp: coords must be a CUDA tensor (no CPU fallback)")
    x = x.detach().to(torch.float32).contiguous()
    n = x.shape[0]
    dev = x.device
    y = torch.empty((n, out_features), dtype=torch.float32, device=dev)
    J = torch.empty((n, out_features, 3), dtype=torch.float32, device=dev) if order >= 1 else None
    H = torch.empty((n, out_features, 3, 3), dtype=torch.float32, device=dev) if order >= 2 else None
    m = _lib.MODES[mode or _JET_MODE["mode"]]
    ws = _jet_workspace(dev) if (order >= 1 and out_features == 1 and m in (_lib.MODE_AUTO, _lib.MODE_TC_REVERSE)) else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().inf3dp_siren_jet_ws(_lib.ptr(blob), _lib.ptr(x), n, order, _lib.ptr(y), _lib.ptr(J),
                                                  _lib.ptr(H), m, _lib.ptr(ws), ws.numel() if ws is not None else 0,
                                                  _lib.stream_ptr(dev)))
    return y, J, H


def value_grad_rows(blob, x, out=None):
    """x [n,3] (CUDA, f32) -> rows [n,4] = (value, d/dx, d/dy, d/dz) from the reverse-mode engine, optionally into `out`."""
    if not x.is_cuda:
        raise RuntimeError("inf3dp: coords must be a CUDA tensor (no CPU fallback)")
    x = x.detach().to(torch.float32).contiguous()
    n = x.shape[0]
    dev = x.device
    if out is None:
        out = torch.empty((n, 4), dtype=torch.float32, device=dev)
    elif (not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or out.numel() < 4 * n):
        raise RuntimeError("inf3dp: `out` must be a contiguous CUDA float32 tensor with room for [n,4] rows")
    ws = _jet_workspace(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().inf3dp_siren_value_grad_rows(_lib.ptr(blob), _lib.ptr(x), n, _lib.ptr(out), _lib.ptr(ws),
                                                           ws.numel(), _lib.stream_ptr(dev)))
    return out.view(-1, 4)[:n] if out.dim() != 2 or out.shape[0] != n else out


def value_grad_rows_peers(blob, x, peer_ptrs, multicast_ptr=0):
    """x [n,3] (CUDA, f32): the (value, grad) rows are stored by the kernel's epilogue directly at the given peer-mapped
    device addresses (ints; one per rank of the job, each = destination of row 0 of this call), or through ONE NVLink
    multicast address -- include/inf3dp.h: inf3dp_siren_value_grad_rows_peers.  Nothing is returned: the rows live in the
    callers' gathered buffers (see inf3dp.dist.FusedRowGather)."""
    if not x.is_cuda:
        raise RuntimeError("inf3dp: coords must be a CUDA tensor (no CPU fallback)")
    x = x.detach().to(torch.float32).contiguous()
    n = x.shape[0]
    dev = x.device
    ws = _jet_workspace(dev)
    arr = (ctypes.c_void_p * max(1, len(peer_ptrs)))(*[int(p) for p in peer_ptrs])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().inf3dp_siren_value_grad_rows_peers(_lib.ptr(blob), _lib.ptr(x), n, arr, len(peer_ptrs),
                                                                 ctypes.c_void_p(int(multicast_ptr)) if multicast_ptr else None,
                                                                 _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))


class _SirenJacobian(torch.autograd.Function):
    """J(x) as a differentiable node: its backward contracts with the Hessian of an order-2 launch."""

    @staticmethod
    def forward(ctx, x, net, J_pre):
        ctx.net = net
        ctx.save_for_backward(x)
        if J_pre is not None:
            return J_pre.view(*x.shape[:-1], net.out_features, 3)
        _, J, _ = net._jet(x.reshape(-1, x.shape[-1]), 1)
        return J.view(*x.shape[:-1], net.out_features, 3)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gJ):
        (x,) = ctx.saved_tensors
        net = ctx.net
        _, _, H = net._jet(x.reshape(-1, x.shape[-1]), 2)
        g = torch.einsum("noi,noik->nk", gJ.reshape(-1, net.out_features, 3).to(H.dtype), H)
        return g.view_as(x), None, None


class _SirenValue(torch.autograd.Function):
    """y(x); backward = grad_out . J with J from the fused forward-mode kernel (twice differentiable)."""

    @staticmethod
    def forward(ctx, x, net, eager):
        ctx.net = net
        flat = x.reshape(-1, x.shape[-1])
        if eager:
            y, J, _ = net._jet(flat, 1)
        else:
            y, _, _ = net._jet(flat, 0)
            J = None
        ctx.save_for_backward(x, J) if J is not None else ctx.save_for_backward(x)
        ctx.has_J = J is not None
        return y.view(*x.shape[:-1], net.out_features)

    @staticmethod
    def backward(ctx, gy):
        if not ctx.needs_input_grad[0]:
            return None, None, None
        saved = ctx.saved_tensors
        x = saved[0]
        J_pre = saved[1] if ctx.has_J else None
        J = _SirenJacobian.apply(x, ctx.net, J_pre)
        gx = (gy.unsqueeze(-1) * J).sum(-2)
        return gx, None, None


class FusedSirenMixin:
    """Shared query plumbing of the drop-in modules (expects .out_features, ._packed, ._weights())."""

    lazy_jacobian = False  # True: forward computes values only and the Jacobian is launched on demand in backward

    # ---- inference only, and loud about it --------------------------------------------------------------------------
    # The fused kernels treat the weights as constants: there is no autograd path from the outputs to the parameters.
    # In a reference training loop loss.backward() would still succeed (the coordinates are a leaf), every parameter's
    # .grad would stay None and optimizer.step() would silently do nothing.  So: parameters are created frozen,
    # train(True) raises, and forward() raises if someone re-enabled requires_grad on a parameter.
    def _freeze(self):
        for p in self.parameters():
            p.requires_grad_(False)

    def train(self, mode: bool = True):
        if mode:
            raise RuntimeError(f"inf3dp.{type(self).__name__} is inference-only: the fused kernels have no gradient path to "
                               "the weights, a training loop on it would silently not train.  Train with the reference's "
                               "modules.SingleBVPNet / MLPClassifier and load the state_dict here (same keys); call .eval() "
                               "on this module (the reference's query helpers do, utils.py:146,228).")
        return super().train(False)

    def _assert_frozen(self):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError(f"inf3dp.{type(self).__name__}: a parameter has requires_grad=True, but gradients only flow to "
                               "the coordinates (inference-only drop-in); weight gradients would silently stay None")

    def _jet(self, flat_x, order, mode=None):
        ws, bs = self._weights()
        blob = self._packed.get(ws, bs)
        return jet(blob, flat_x, order, self.out_features, mode)

    @torch.no_grad()
    def query(self, coords, order=1, mode=None):
        """Fused value / Jacobian / Hessian query without autograd bookkeeping.

        coords [..., in] -> (y [..., out], J [..., out, 3] or None, H [..., out, 3, 3] or None)."""
        flat = coords.reshape(-1, coords.shape[-1])
        y, J, H = self._jet(flat, order, mode)
        lead = coords.shape[:-1]
        y = y.view(*lead, self.out_features)
        if J is not None:
            J = J.view(*lead, self.out_features, 3)
        if H is not None:
            H = H.view(*lead, self.out_features, 3, 3)
        return y, J, H


    @torch.no_grad()
    def query_rows(self, coords, out=None):
        """Value + gradient as packed rows [n,4] = (value, d/dx, d/dy, d/dz) (single-output sine field networks): the
        layout of the reference's host result (utils.py:249-253) and of the multi-GPU all-gather."""
        ws, bs = self._weights()
        blob = self._packed.get(ws, bs)
        return value_grad_rows(blob, coords.reshape(-1, coords.shape[-1]), out)


    @torch.no_grad()
    def query_rows_peers(self, coords, peer_ptrs, multicast_ptr=0):
        """query_rows with the rows stored straight into peer-mapped buffers / an NVLink multicast address."""
        ws, bs = self._weights()
        blob = self._packed.get(ws, bs)
        value_grad_rows_peers(blob, coords.reshape(-1, coords.shape[-1]), peer_ptrs, multicast_ptr)


class SingleBVPNet(FusedSirenMixin, nn.Module):
    """Drop-in for modules.SingleBVPNet (modules.py:119-166), mode='mlp', type in {'sine','relu'}."""

    def __init__(self, out_features=1, type="sine", in_features=2, mode="mlp", hidden_features=256,
                 num_hidden_layers=3, **kwargs):
        super().__init__()
        if mode != "mlp":
            raise NotImplementedError("inf3dp.SingleBVPNet: only mode='mlp' is on the accelerated path "
                                      "(rbf / nerf encodings are unused by INF-3DP, SURVEY.md section 2 #13)")
        if type not in ("sine", "relu"):
            raise NotImplementedError(f"inf3dp.SingleBVPNet: nonlinearity {type!r} is not on the accelerated path")
        self.mode = mode
        self.type = type
        self.in_features, self.out_features = int(in_features), int(out_features)
        self.hidden_features, self.num_hidden_layers = int(hidden_features), int(num_hidden_layers)
        # module tree chosen so that state_dict keys equal the reference's: net.net.{i}.0.{weight,bias}
        dims = [in_features] + [hidden_features] * (num_hidden_layers + 1) + [out_features]
        layers = [nn.Sequential(nn.Linear(dims[i], dims[i + 1])) for i in range(len(dims) - 1)]
        holder = nn.Module()
        holder.net = nn.Sequential(*layers)
        self.net = holder
        # same RNG call order as FCBlock.__init__ (modules.py:66-87): default nn.Linear init per layer, then the
        # nonlinearity's init over all layers, then the first-layer init
        with torch.no_grad():
            if type == "sine":
                for seq in layers:  # sine_init, modules.py:622-627
                    w = seq[0].weight
                    bound = math.sqrt(6 / w.size(-1)) / 30
                    w.uniform_(-bound, bound)
                w0 = layers[0][0].weight  # first_layer_sine_init, modules.py:630-635
                w0.uniform_(-1 / w0.size(-1), 1 / w0.size(-1))
            else:
                for seq in layers:  # init_weights_normal, modules.py:589-593
                    nn.init.kaiming_normal_(seq[0].weight, a=0.0, nonlinearity="relu", mode="fan_in")
        act = _lib.ACT_SINE if type == "sine" else _lib.ACT_RELU
        self._packed = PackedWeights(in_features, hidden_features, num_hidden_layers, out_features, act, 30.0)
        self._freeze()
        nn.Module.train(self, False)

    def _weights(self):
        lins = [seq[0] for seq in self.net.net]
        return [l.weight for l in lins], [l.bias for l in lins]

    def forward(self, model_input, params=None):
        if params is not None:
            raise NotImplementedError("inf3dp.SingleBVPNet: hypernetwork `params` are not supported (inference path)")
        self._assert_frozen()
        coords_org = model_input["coords"].clone().detach().requires_grad_(True)  # modules.py:148
        eager = torch.is_grad_enabled() and not self.lazy_jacobian
        out = _SirenValue.apply(coords_org, self, eager)
        return {"model_in": coords_org, "model_out": out}


class MLPClassifier(FusedSirenMixin, nn.Module):
    """Drop-in for exp_scripts/train_levels.py:35-51 (4 -> 256 -> 256 -> C ReLU logits), keys network.{0,2,4}.*"""

    def __init__(self, input_dim=4, hidden_dim=256, num_classes=3):
        super().__init__()
        self.network = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim),
                                     nn.ReLU(), nn.Linear(hidden_dim, num_classes))
        self.out_features = int(num_classes)
        self._packed = PackedWeights(input_dim, hidden_dim, 1, num_classes, _lib.ACT_RELU, 1.0)
        self._freeze()
        nn.Module.train(self, False)

    def _weights(self):
        lins = [self.network[0], self.network[2], self.network[4]]
        return [l.weight for l in lins], [l.bias for l in lins]

    def forward(self, x):
        self._assert_frozen()
        flat = x.reshape(-1, x.shape[-1])
        y, _, _ = self._jet(flat, 0)
        return y.view(*x.shape[:-1], self.out_features)
