"""Batched query helpers: fused equivalents of the reference's chunking loops (utils.py:140-303).

The reference helpers stay usable unmodified on top of inf3dp.SingleBVPNet; these versions remove their
per-chunk autograd bookkeeping and PCIe ping-pong:
  * field_querying / field_query_with_grads (utils.py:140-176, :217-266): one fused launch per chunk, results stay
    on the device unless the input lives on the host
  * host inputs are streamed: pinned staging buffers, H2D of chunk i+1 and D2H of chunk i-1 overlap the kernel of
    chunk i on separate CUDA streams
"""
from __future__ import annotations

import torch


def _net_of(decoder):
    """Accept an inf3dp network or a reference-style decoder wrapper holding one as `.model`."""
    for cand in (decoder, getattr(decoder, "model", None)):
        if cand is not None and hasattr(cand, "query"):
            return cand
    raise RuntimeError("inf3dp.query: decoder is not (and does not wrap as .model) an inf3dp network")


@torch.no_grad()
def field_querying(decoder, samples, max_batch=1 << 22, mode=None):
    """samples [N,3] on the device -> fields [N] (out=1) or [N,out] on the device."""
    net = _net_of(decoder)
    n = samples.shape[0]
    outs = []
    for h in range(0, n, max_batch):
        y, _, _ = net.query(samples[h:h + max_batch], order=0, mode=mode)
        outs.append(y)
    y = torch.cat(outs, 0) if len(outs) != 1 else outs[0]
    return y.squeeze(-1) if net.out_features == 1 else y


@torch.no_grad()
def field_query_with_grads(decoder, coords, max_batch=1 << 22, mode=None):
    """coords [N,3] on the device -> (fields [N], grads [N,3]) on the device (utils.py:217-266 semantics, out=1)."""
    net = _net_of(decoder)
    n = coords.shape[0]
    fields = torch.empty(n, dtype=torch.float32, device=coords.device)
    grads = torch.empty((n, 3), dtype=torch.float32, device=coords.device)
    for h in range(0, n, max_batch):
        y, J, _ = net.query(coords[h:h + max_batch], order=1, mode=mode)
        fields[h:h + max_batch] = y[:, 0]
        grads[h:h + max_batch] = J[:, 0, :]
    return fields, grads


def gen_voxel_queries(N, cube_size=1.0, device=None):
    """The reference's query grid (sdf_meshing.py:46-65) built on `device`, bit-identical to the reference's CPU
    result INCLUDING its true-division shear: y and x indices are the float32 values (i / N) % N and
    ((i / N) / N) % N, not integers.  Returns (samples [N^3, 3] f32, origin, voxel_size)."""
    voxel_size = cube_size * 2.0 / (N - 1)
    idx = torch.arange(0, N ** 3, 1, dtype=torch.long, device=device)
    samples = torch.zeros(N ** 3, 3, device=device)
    samples[:, 2] = idx % N
    q1 = idx / N            # true division on a LongTensor -> float32, as in the reference
    samples[:, 1] = q1 % N
    samples[:, 0] = (q1 / N) % N
    del idx, q1
    samples[:, 0] = (samples[:, 0] * voxel_size) + (-cube_size)
    samples[:, 1] = (samples[:, 1] * voxel_size) + (-cube_size)
    samples[:, 2] = (samples[:, 2] * voxel_size) + (-cube_size)
    return samples, (-cube_size, -cube_size, -cube_size), voxel_size


class HostPipeline:
    """Value+gradient queries for HOST-resident coordinates (the reference's calling pattern: CPU tensors in,
    CPU tensors out, utils.py:232-253) with copies overlapped with compute.

    Three CUDA streams (H2D, compute, D2H) and double-buffered device + pinned host staging; per chunk the
    timeline is  H2D(i+1) || jet(i) || D2H(i-1).
    """

    def __init__(self, decoder, chunk=1 << 22, device=None, mode=None):
        self.net = _net_of(decoder)
        self.chunk = int(chunk)
        self.mode = mode
        self.rows_engine = mode in (None, "auto", "tc_reverse") and hasattr(self.net, "query_rows") and \
            getattr(self.net, "out_features", 0) == 1 and getattr(self.net, "in_features", 0) == 3 and \
            getattr(self.net, "type", "") == "sine" and getattr(self.net, "hidden_features", 0) == 256 and \
            getattr(self.net, "num_hidden_layers", 0) == 3
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("inf3dp.HostPipeline needs a CUDA device")
        dev = self.device
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.d_x = [torch.empty((self.chunk, 3), dtype=torch.float32, device=dev) for _ in range(2)]
        self.d_o = [torch.empty((self.chunk, 4), dtype=torch.float32, device=dev) for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_k = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]

    @torch.no_grad()
    def run(self, coords_host, out_host):
        """coords_host [N,3] pinned f32, out_host [N,4] pinned f32 (value, dx, dy, dz).  Returns after the last
        D2H has completed (the caller owns valid host results on return)."""
        n = coords_host.shape[0]
        cur = torch.cuda.current_stream(self.device)
        nchunks = (n + self.chunk - 1) // self.chunk
        for i in range(nchunks):
            b = i & 1
            lo, hi = i * self.chunk, min(n, (i + 1) * self.chunk)
            m = hi - lo
            with torch.cuda.stream(self.s_in):
                if i >= 2:
                    self.s_in.wait_event(self.ev_k[b])  # the kernel that read this staging buffer is done
                self.d_x[b][:m].copy_(coords_host[lo:hi], non_blocking=True)
                self.ev_in[b].record(self.s_in)
            cur.wait_event(self.ev_in[b])
            if i >= 2:
                cur.wait_event(self.ev_out[b])  # the D2H that read this output buffer is done
            if self.rows_engine:   # the reverse-mode engine writes the (value, grad) rows in place
                self.net.query_rows(self.d_x[b][:m], out=self.d_o[b][:m])
            else:
                y, J, _ = self.net.query(self.d_x[b][:m], order=1, mode=self.mode)
                self.d_o[b][:m, 0:1].copy_(y)
                self.d_o[b][:m, 1:4].copy_(J[:, 0, :])
            self.ev_k[b].record(cur)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_k[b])
                out_host[lo:hi].copy_(self.d_o[b][:m], non_blocking=True)
                self.ev_out[b].record(self.s_out)
        self.s_out.synchronize()
        return out_host
