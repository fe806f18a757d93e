"""ORACLE (test infrastructure, not product code) -- torch-CPU port of the reference's SIREN query path.

This is the `cpu_baseline` / `--impl reference` vehicle of bench.py and a second checker for tests: the same
operator sequence the reference executes (reference files relative to /root/reference):
  * BatchLinear.forward  modules.py:16-25   -> input.matmul(weight^T); output += bias
  * Sine.forward         modules.py:32-34   -> torch.sin(30 * input)
  * SingleBVPNet.forward modules.py:143-160 -> coords.clone().detach().requires_grad_(True); {'model_in','model_out'}
  * gradient             diff_operators.py:128-132 -> torch.autograd.grad(y, [x], ones, create_graph=True)
  * hessian3d            diff_operators.py:27-44
  * field_query_with_grads chunk loop  utils.py:217-266 (max_batch chunks, results stored per chunk)
/root/reference is a Python reference and cannot travel to the GPU box, hence this port (kind = "port").
It is validated against the imported reference in tests/golden/make_golden.py (bitwise equal on CPU).
"""
from __future__ import annotations

import torch


class PortSiren(torch.nn.Module):
    def __init__(self, layers):
        """layers: [(W [out,in], b [out])] float32 arrays/tensors; last layer linear."""
        super().__init__()
        self.W = torch.nn.ParameterList([torch.nn.Parameter(torch.as_tensor(W, dtype=torch.float32).clone()) for W, _ in layers])
        self.b = torch.nn.ParameterList([torch.nn.Parameter(torch.as_tensor(b, dtype=torch.float32).clone()) for _, b in layers])

    def forward(self, model_input):
        coords = model_input["coords"].clone().detach().requires_grad_(True)
        h = coords
        n = len(self.W)
        for i in range(n):
            h = h.matmul(self.W[i].permute(-1, -2))
            h += self.b[i].unsqueeze(-2)
            if i != n - 1:
                h = torch.sin(30 * h)
        return {"model_in": coords, "model_out": h}


def gradient(y, x, grad_outputs=None):
    if grad_outputs is None:
        grad_outputs = torch.ones_like(y)
    return torch.autograd.grad(y, [x], grad_outputs=grad_outputs, create_graph=True)[0]


def hessian3d(y, x):
    n, ch = y.shape[0], y.shape[1]
    gy = torch.ones_like(y[..., 0])
    h = torch.zeros(n, ch, x.shape[-1], x.shape[-1])
    for i in range(ch):
        dydx = torch.autograd.grad(y[..., i], x, gy, create_graph=True)[0]
        for j in range(x.shape[-1]):
            h[..., i, j, :] = torch.autograd.grad(dydx[..., j], x, gy, create_graph=True)[0]
    return h


def field_query_with_grads(model, coords, max_batch=32 ** 3):
    """utils.py:217-266 minus the .cuda() hops: chunked value + gradient into preallocated CPU tensors."""
    samples = coords.float()
    n = samples.shape[0]
    fields = torch.zeros(n)
    grads = torch.zeros(n, 3)
    head = 0
    while head < n:
        sub = samples[head:min(head + max_batch, n), :]
        with torch.enable_grad():
            out = model({"coords": sub})
            fields[head:head + sub.shape[0]] = out["model_out"].squeeze(-1).detach()
            grads[head:head + sub.shape[0]] = gradient(out["model_out"], out["model_in"]).detach()
        head += max_batch
    return fields, grads
