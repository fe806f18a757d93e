#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native INF-3DP hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

Metric (BASELINE.json): SIREN value+grad points/s.  Workload at every N: BASELINE.json configs[1], the dense
512^3 grid evaluation with normals of the SDF network (sdf_meshing path): coordinates = the reference's
gen_voxle_queries(512) (134 217 728 points, shear quirk included), network = sdf_config arch
(3->256->256->256->256->1, sine, random init seed 2048), outputs = field value, gradient and unit normal per point.
A step is one pass over the whole grid.  With N ranks every rank evaluates its own 512^3 grid shard of an
N*512^3-point job (weak scaling) and the (value, gradient) rows are all-gathered: the query kernel's epilogue stores
every row straight into all ranks' gathered buffers over NVLink (multicast when the fabric has it, else peer stores) --
the collective is fused into the kernel, nothing follows it but a cross-rank barrier (inf3dp.dist.FusedRowGather).

`value`  : whole-job points/s with coordinates resident in HBM (CUDA events, max over ranks).
`e2e`    : the same metric through the public host-buffer API: pinned host coordinates -> H2D -> fused kernel (+ fused
           all-gather at N > 1) -> D2H of (value, gradient) rows into pinned host memory, copies inside the timed region.
           N = 1: inf3dp.HostPipeline; N > 1: inf3dp.dist.ShardedHostPipeline, every rank copying out the rows of its
           NEIGHBOUR's shard, so that the collective lies on the timed path.  The host result is checked against a
           device-resident evaluation and the run aborts on a mismatch.
`roofline`: dominant kernel (siren_rev_kernel, the reverse-mode tcgen05 engine): algorithmic FLOPs of the formulation it
           executes (790 528 per point: one forward + one backward sweep, SURVEY.md 8d "reverse-mode equivalent")
           / launch duration against the measured dense tensor peak in MEASURED_PEAKS.json.  INF3DP_ENGINE=tc_precise
           selects the forward-mode jet engine instead (1 576 448 FLOP per point, the figure it is then judged on).
`cpu_baseline`: the reference's own torch path on the host cores (staged unmodified under oracle/_ref/py: kind
           "reference"; the torch port of its operator sequence otherwise: kind "port"), bounded sample.
`--impl reference`: that CPU path alone, same metric, bounded sample per step.
Other BASELINE.json configurations ride in the same line: `isocontour` (second metric, layers/s; dense contract, compact
rows, e2e of both, the compiled reference extension timed on the same GPU), `intersection` (a6 at the reference's size),
`hessian_sweep` (cfg 5), `collision_sweep` (cfg 4), `volume` (N4).  At N > 1 those shard the way SURVEY.md 8e says
(isovalues / slice isovalues / points / waypoints) with their NCCL all-gather, strong scaling, checked against N = 1.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "inf-3dp_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

FLOP_FORWARD_MODE = 1_576_448       # forward-mode value+gradient (value + 3 tangent columns), SURVEY.md section 8d
FLOP_REVERSE_MODE = 790_528         # one forward + one backward sweep (what the reference's autograd executes), ibid.
ENGINES = {  # engine -> (kernel, algorithmic FLOP/point, tensor FLOP/point actually issued, dtype note)
    "tc_reverse": ("siren_rev_kernel<CTAS=2>", FLOP_REVERSE_MODE, 3 * 6 * 2 * 256 * 256, "f16x3-split operands, f32 accumulate"),
    "tc_precise": ("siren_tc_kernel<ORDER=1,PRECISE,OUTC=1>", FLOP_FORWARD_MODE, 3 * 1_572_864, "f16x3-split operands, f32 accumulate"),
    "tc_fast": ("siren_tc_kernel<ORDER=1,FAST,OUTC=1>", FLOP_FORWARD_MODE, 1_572_864, "f16 operands, f32 accumulate (NOT parity grade)"),
    "simt": ("siren_simt_kernel<4>", FLOP_FORWARD_MODE, 0, "f32"),
}
GRID_N = 512
WORKLOAD = "bunny SDF dense 512^3 grid evaluation (sdf_meshing path) with normals"
CPU_SAMPLE = 262144


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm": float(d["hbm_gbs"]), "tf_burst": float(d["bf16_tflops"]),
                "tf_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


def measured_traffic(key, source_file):
    """DRAM bytes per launch of a kernel from a committed ncu capture (profiles/traffic.json), valid only while the kernel's
    source file is byte-identical to the one that was profiled; otherwise None (never a stale constant)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        rec = json.load(open(path))[key]
        sha = hashlib.sha256(open(os.path.join(ROOT, "inf-3dp_b200", "csrc", source_file), "rb").read()).hexdigest()
        return float(rec["dram_bytes"]) if rec.get("source_sha256") == sha else None
    except Exception:
        return None


def golden_layers():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "siren_golden.npz"))
    return [(g[f"sdf/net.net.{i}.0.weight"], g[f"sdf/net.net.{i}.0.bias"]) for i in range(5)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1]); power.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort(); power.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w": power[len(power) // 2] if power else None}


def grid_slab_cpu(N, start, count, cube_size=1.0):
    """Rows [start, start + count) of the reference's gen_voxle_queries(N) (sdf_meshing.py:46-65) on the CPU, same float32
    operations (true division on a LongTensor included), without building the N^3 grid."""
    import torch
    voxel_size = cube_size * 2.0 / (N - 1)
    idx = torch.arange(start, start + count, dtype=torch.long)
    s = torch.zeros(count, 3)
    s[:, 2] = idx % N
    q1 = idx / N
    s[:, 1] = q1 % N
    s[:, 0] = (q1 / N) % N
    for c in range(3):
        s[:, c] = (s[:, c] * voxel_size) + (-cube_size)
    return s


_CPU_REF = {}


def cpu_reference_points_per_s(sample_points, repeats, threads):
    """The reference's own torch path on the host cores: utils.field_query_with_grads (utils.py:217-266, max_batch 32^3) on
    a slab of the 512^3 grid -- the UNMODIFIED reference staged under oracle/_ref/py (Tensor.cuda shimmed to the identity)
    when it is there (kind "reference"), otherwise the torch port of its operator sequence (kind "port")."""
    import numpy as np
    import torch
    torch.set_num_threads(threads)
    start = (GRID_N ** 3) // 2            # a slab from the middle of the grid
    coords = grid_slab_cpu(GRID_N, start, sample_points)
    if "fn" not in _CPU_REF:
        from oracle import build_ref
        layers = golden_layers()
        if build_ref.python_staged():
            import contextlib
            import io
            modules, _, utils = build_ref.load_python(cpu_shim=True)
            with contextlib.redirect_stdout(io.StringIO()):
                model = modules.SingleBVPNet(type="sine", in_features=3)
            model.load_state_dict({f"net.net.{i}.0.{k}": torch.from_numpy(np.asarray(v)) for i, (W, b) in enumerate(layers)
                                   for k, v in (("weight", W), ("bias", b))})

            class Decoder(torch.nn.Module):
                def __init__(self, m):
                    super().__init__()
                    self.model = m

                def forward(self, c):
                    return self.model({"coords": c})
            dec = Decoder(model)
            _CPU_REF["fn"] = lambda c: utils.field_query_with_grads(dec, c, max_batch=32 ** 3, no_print=True)
            _CPU_REF["kind"] = "reference"
            _CPU_REF["what"] = "the unmodified reference utils.field_query_with_grads + modules.SingleBVPNet (oracle/_ref/py)"
        else:
            from oracle import siren_torch_port as port
            model = port.PortSiren([(W.astype(np.float32), b.astype(np.float32)) for W, b in layers])
            model.eval()
            _CPU_REF["fn"] = lambda c: port.field_query_with_grads(model, c, max_batch=32 ** 3)
            _CPU_REF["kind"] = "port"
            _CPU_REF["what"] = "oracle/siren_torch_port.field_query_with_grads (torch port of the reference's operator sequence)"
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        _CPU_REF["fn"](coords.clone())
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample_points / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    times = []
    import torch
    for i in range(args.warmup + args.steps):
        _, dt = cpu_reference_points_per_s(CPU_SAMPLE, 1, cores)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = CPU_SAMPLE / (ms / 1e3)
    desc = (f"{CPU_SAMPLE} consecutive points from the middle of the 512^3 grid per step, {_CPU_REF['what']}, max_batch 32^3, "
            f"{cores} host threads")
    line = {"impl": "reference", "metric": "siren_value_grad_points_per_s", "value": value, "unit": "points/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "network": "3-256-256-256-256-1 sine, seed 2048", "grid": GRID_N},
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": _CPU_REF["kind"], "sample": desc},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "torch": torch.__version__}
    emit_json_line(line)


# ---------------------------------------------------------------------------------------------------------------------
def _timeit(fn, dev, warm, iters):
    import torch
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / iters, out


def _max_over_ranks(ms, dev, world):
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def _contour_field(V):
    import numpy as np
    return (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)


def bench_isocontours(dev, peaks, rank, world, extras):
    """BASELINE.json configs[2]: 200-isovalue isocontour extraction on a closed synthetic mesh (SURVEY.md 8d cfg 3).
    N > 1: the isovalues are sharded (K / N per rank, mesh and field replicated) and the dense [K/N, F, 3] slabs and counts are
    all-gathered -- strong scaling, gathered result compared with the single-GPU one."""
    import numpy as np
    import torch
    import inf3dp
    from inf3dp.testing import torus_mesh
    from inf3dp.dist import sharded_isocontours
    V, F = torus_mesh(1000, 500)           # 1.0 M faces, 0.5 M vertices
    f = _contour_field(V)
    K = 200
    iso = np.linspace(f.min(), f.max(), K).astype(np.float32)
    t = [torch.from_numpy(a).to(dev) for a in (V, F, f, iso)]
    L = inf3dp._lib.lib()
    nV, nF = len(V), len(F)
    merged = torch.empty((K, nF, 3), dtype=torch.float32, device=dev)
    nums = torch.empty(K, dtype=torch.int32, device=dev)
    wsb = L.inf3dp_extract_isocontours_workspace_bytes(nV, nF, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = inf3dp._lib.stream_ptr(dev)
    P = inf3dp._lib.ptr

    def call():   # the C ABI call of the drop-in, device resident
        inf3dp._lib.check(L.inf3dp_extract_isocontours(*[P(a) for a in t], nV, nF, K, P(merged), P(nums), P(ws), wsb, st))
    ms_single, _ = _timeit(call, dev, 3, 10)
    seg = (inf3dp._lib.ctypes.c_int32 * K)()
    inf3dp._lib.check(L.inf3dp_extract_isocontours_status(P(ws), K, seg, st))
    S = int(sum(seg)); C = int(nums.sum().item())
    contract = 12 * nF + 16 * nV + 4 * K + 12 * K * nF + 4 * K
    useful = 12 * nF + 16 * nV + 4 * K + 12 * (S + C) + 4 * K
    out = {"metric": "isocontour_layers_per_s", "unit": "layers/s",
           "config": {"workload": "200-isovalue isocontour extraction, closed torus mesh", "faces": nF, "vertices": nV,
                      "isovalues": K, "segments": S, "loops": C}}
    if world > 1:
        ms_sh, (gm, gn) = _timeit(lambda: sharded_isocontours(*t), dev, 2, 5)
        ms_sh = _max_over_ranks(ms_sh, dev, world)
        if not (torch.equal(gm, merged) and torch.equal(gn, nums)):
            raise SystemExit(f"bench.py: rank {rank}: isovalue-sharded isocontours differ from the single-GPU result")
        del gm, gn
        # the exchange that makes the split pay: compact row streams (used rows only) instead of the dense contract
        from inf3dp.dist import sharded_isocontours_compact
        ms_c, (rows, offs, counts, cnums) = _timeit(lambda: sharded_isocontours_compact(*t), dev, 2, 5)
        ms_c = _max_over_ranks(ms_c, dev, world)
        k_chk = (rank * 37) % K
        u = int(counts[k_chk])
        if not (torch.equal(cnums, nums) and torch.equal(rows[int(offs[k_chk]): int(offs[k_chk]) + u], merged[k_chk, :u])):
            raise SystemExit(f"bench.py: rank {rank}: isovalue-sharded compact rows differ from the single-GPU dense contract")
        ms_c1, _ = _timeit(lambda: inf3dp.extract_isocontours_compact(*t), dev, 2, 5)
        out.update({"value": K / (ms_sh / 1e3), "ms_per_call": ms_sh, "n_gpus": world,
                    "scaling": f"strong: {K} isovalues sharded over {world} ranks, NCCL all-gather of the dense [K/N,F,3] slabs "
                               "and counts inside the timed region, result identical to N = 1 (the 2.4 GB dense contract makes the "
                               "gather cost more than the extraction: see `compact_sharded`)",
                    "single_gpu_ms_per_call": ms_single,
                    "compact_sharded": {"api": "inf3dp.dist.sharded_isocontours_compact (K/N isovalues per rank, all-gather of the used rows only)",
                                        "ms_per_call": ms_c, "layers_per_s": K / (ms_c / 1e3), "single_gpu_ms_per_call": ms_c1,
                                        "rows_gathered": int(rows.shape[0])}})
        return out
    # ---- N = 1: dense contract roofline, compact path, end to end, comparators --------------------------------------------------
    out.update({"value": K / (ms_single / 1e3), "ms_per_call": ms_single, "gpu_launches_per_call": 9,
                "roofline": {"bound": "hbm", "achieved": contract / (ms_single / 1e3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                             "frac": contract / (ms_single / 1e3) / 1e9 / peaks["hbm"],
                             "traffic": measured_traffic("contours_dense", "contours.cu"), "traffic_unit": "DRAM bytes per call (ncu, profiles/traffic.json)",
                             "bytes_contract": contract, "bytes_useful": useful,
                             "useful_gbs": useful / (ms_single / 1e3) / 1e9, "peak_source": peaks["source"],
                             "note": "99 % of the contract bytes are the compulsory zero fill of the dense [K,F,3] result"}})
    # compact rows (N2): what a caller who does not need the dense contract pays, judged on useful bytes
    ms_compact, (rows, offs, counts, cnums) = _timeit(lambda: inf3dp.extract_isocontours_compact(*t), dev, 2, 5)
    total_rows = int((offs + counts).max().item())
    compact_bytes = 12 * nF + 16 * nV + 4 * K + 12 * total_rows + 16 * K
    out["compact"] = {"api": "inf3dp.extract_isocontours_compact (CSR rows, no dense [K,F,3] tensor; includes its status read-back)",
                      "ms_per_call": ms_compact, "layers_per_s": K / (ms_compact / 1e3), "rows": total_rows,
                      "bytes_useful": compact_bytes, "useful_gbs": compact_bytes / (ms_compact / 1e3) / 1e9,
                      "frac_of_hbm_peak": compact_bytes / (ms_compact / 1e3) / 1e9 / peaks["hbm"],
                      "note": "latency / launch bound: 9 dependent launches over ~26 MB of useful bytes"}
    # end to end through the drop-in API with HOST inputs and outputs (what toolpath_utils.py:23-52 does around the op)
    host_in = [torch.from_numpy(a).pin_memory() for a in (V, F, f, iso)]
    dense_host = torch.empty((K, nF, 3), dtype=torch.float32, pin_memory=True)
    nums_host = torch.empty(K, dtype=torch.int32, pin_memory=True)

    def e2e_dense():
        td = [a.to(dev, non_blocking=True) for a in host_in]
        m, n = inf3dp.extract_isocontours(*td)
        dense_host.copy_(m, non_blocking=True); nums_host.copy_(n, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    t0 = time.perf_counter(); e2e_dense(); e2e_dense(); torch.cuda.synchronize(dev)
    t0 = time.perf_counter(); e2e_dense(); ms_e2e_dense = (time.perf_counter() - t0) * 1e3
    if not (torch.equal(dense_host, merged.cpu()) and torch.equal(nums_host, nums.cpu())):
        raise SystemExit("bench.py: isocontour e2e (dense) result differs from the device-resident call")
    rows_host = torch.empty((total_rows, 3), dtype=torch.float32, pin_memory=True)

    def e2e_compact():
        td = [a.to(dev, non_blocking=True) for a in host_in]
        r, o, c, n = inf3dp.extract_isocontours_compact(*td)
        rows_host.copy_(r[:total_rows], non_blocking=True)
        oh, ch, nh = o.cpu(), c.cpu(), n.cpu()
        torch.cuda.current_stream(dev).synchronize()
        return oh, ch, nh
    e2e_compact(); e2e_compact()
    t0 = time.perf_counter(); oh, ch, nh = e2e_compact(); ms_e2e_compact = (time.perf_counter() - t0) * 1e3
    k_chk = K // 2
    u = int(ch[k_chk])
    if not torch.equal(rows_host[int(oh[k_chk]): int(oh[k_chk]) + u], merged[k_chk, :u].cpu()):
        raise SystemExit("bench.py: isocontour e2e (compact) rows differ from the dense contract's used prefix")
    h2d = sum(a.numel() * a.element_size() for a in host_in)
    out["e2e"] = {"dense": {"value": K / (ms_e2e_dense / 1e3), "unit": "layers/s", "ms_per_call": ms_e2e_dense,
                            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": dense_host.numel() * 4 + K * 4,
                            "api": "inf3dp.extract_isocontours on host inputs, dense [K,F,3] + counts copied to pinned host memory"},
                  "compact": {"value": K / (ms_e2e_compact / 1e3), "unit": "layers/s", "ms_per_call": ms_e2e_compact,
                              "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": total_rows * 12 + K * 16,
                              "api": "inf3dp.extract_isocontours_compact on host inputs, used rows + offsets/counts copied to pinned host memory"}}
    del dense_host, rows_host
    if extras:
        # reported CPU baseline: the single-thread C restatement of the reference kernels on a bounded sample (the reference
        # has no CPU implementation of this op); checked against the GPU rows while at it
        from oracle import contour_oracle as co
        sel = np.ascontiguousarray(iso[::25][:8])
        t0 = time.perf_counter()
        mo, no, so_ = co.extract_isocontours(V, F, f, sel)
        dt = time.perf_counter() - t0
        rws = differ = 0
        worst = 0.0
        for j, k in enumerate(range(0, K, 25)[:8]):     # same piece count => same row order: compare row by row
            if int(nums[k]) == int(no[j]):
                mine = merged[k].cpu().numpy()
                used = int(so_[j]) + int(no[j])
                d = np.abs(mine[:used] - mo[j][:used]).max(axis=1) if used else np.zeros(0)
                rws += used
                differ += int((d > 0).sum())
                worst = max(worst, float(d.max()) if used else 0.0)
        out["cpu_baseline"] = {"value": len(sel) / dt, "unit": "layers/s", "cores": 1, "kind": "port",
                               "sample": f"{len(sel)} of the {K} isovalues, oracle/contour_oracle.c (gcc -O2, one thread)",
                               # segments shorter than the reference's 1e-6 merge radius can be entered through either end (DESIGN.md divergence 1)
                               "check_vs_gpu": {"rows_compared": rws, "rows_not_bit_identical": differ, "max_abs_vertex_diff": worst}}
        # the GPU comparator BASELINE.md 5.6(b) names: the reference's own contour_cuda extension on this GPU, same inputs
        from oracle import build_ref
        if os.path.exists(build_ref.so_path()):
            ref = build_ref.load()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            rm, rn = ref.extract_isocontours(*t)
            torch.cuda.synchronize(dev)
            ms_ref = (time.perf_counter() - t0) * 1e3
            same_counts = int((rn == nums).sum().item())
            out["gpu_reference"] = {"what": "the reference's compiled contour_cuda extract_isocontours (oracle/_ref, isocontours.cu:216-293) "
                                            "on the same B200, same inputs, one call",
                                    "ms_per_call": ms_ref, "layers_per_s": K / (ms_ref / 1e3), "speedup_of_ours": ms_ref / ms_single,
                                    "isovalues_with_identical_loop_count": same_counts}
            del rm, rn
    return out


def bench_intersection(dev, peaks, rank, world, extras):
    """a6 at the reference's own size (slice_toolpath.py:125): corner-expanded [255,255,255,8] fields x 3, Ns = 500, N1 = N2 = 128,
    through the drop-in entry point (contract layout).  N > 1: the slice isovalues are sharded (output dim 0), the grid is
    replicated, the three outputs are all-gathered on dim 0."""
    import numpy as np
    import torch
    import inf3dp
    from inf3dp.dist import sharded_fields_intersection
    from inf3dp.testing import CORNER_OFFSETS
    D, Ns, N1, N2 = 255, 500, 128, 128
    lin = torch.linspace(-1, 1, D + 1, device=dev)
    gx, gy, gz = torch.meshgrid(lin, lin, lin, indexing="ij")
    base = [gz + 0.1 * torch.sin(3 * gx), gx + 0.2 * gy ** 2, gy - 0.15 * gz * gx]
    del gx, gy, gz
    fields = [torch.stack([b[o[0]:o[0] + D, o[1]:o[1] + D, o[2]:o[2] + D] for o in CORNER_OFFSETS], -1).contiguous() for b in base]
    idx = torch.stack(torch.meshgrid(*[torch.arange(D, device=dev, dtype=torch.float32)] * 3, indexing="ij"), -1)
    centers = ((idx + 0.5) * (2.0 / D) - 1.0).contiguous()
    del idx, base
    isos = [torch.linspace(-0.9, 0.9, Ns, device=dev), torch.linspace(-0.8, 0.8, N1, device=dev), torch.linspace(-0.8, 0.8, N2, device=dev)]
    full = lambda: inf3dp.get_fields_intersection_inter_slice(*fields, *isos, centers, sync=False)
    ms_single, (m0, i0, p0) = _timeit(full, dev, 2, 5)
    B = 108 * D ** 3 + 4 * (Ns + N1 + N2) + 19 * Ns * N1 * N2
    out = {"metric": "intersection_ms", "config": {"workload": "get_fields_intersection_inter_slice at the reference's size "
                                                                "(slice_toolpath.py:125)", "D": D, "isovalues": [Ns, N1, N2]},
           "cells_hit": int(m0.sum().item()), "launches_per_call": 4}
    if world > 1:
        sharded = lambda: sharded_fields_intersection(*fields, *isos, centers)     # local slice index -> global, 3 all-gathers
        ms_sh, (gm, gi, gp) = _timeit(sharded, dev, 2, 5)
        ms_sh = _max_over_ranks(ms_sh, dev, world)
        if not (torch.equal(gm, m0) and torch.equal(gi, i0) and torch.equal(gp.nan_to_num(), p0.nan_to_num())):
            raise SystemExit(f"bench.py: rank {rank}: slice-isovalue-sharded intersection differs from the single-GPU result")
        out.update({"ms_per_call": ms_sh, "n_gpus": world, "single_gpu_ms_per_call": ms_single,
                    "scaling": f"strong: {Ns} slice isovalues sharded over {world} ranks, NCCL all-gather of the three outputs inside "
                               "the timed region, result identical to N = 1"})
        return out
    out.update({"ms_per_call": ms_single,
                "roofline": {"bound": "hbm", "achieved": B / (ms_single / 1e3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                             "frac": B / (ms_single / 1e3) / 1e9 / peaks["hbm"], "bytes_algorithmic": B,
                             "traffic": measured_traffic("intersection", "intersection.cu"),
                             "traffic_unit": "DRAM bytes per call (ncu, profiles/traffic.json)", "peak_source": peaks["source"],
                             "formula": "108 D^3 + 4 (Ns + N1 + N2) + 19 Ns N1 N2 (SURVEY.md 8d)"}})
    if extras:
        from oracle import build_ref
        if os.path.exists(build_ref.so_path()):
            ref = build_ref.load()
            call = lambda: ref.get_fields_intersection_inter_slice(*fields, *isos, centers)
            call(); torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            rm, ri, rp = call()
            torch.cuda.synchronize(dev)
            ms_ref = (time.perf_counter() - t0) * 1e3
            out["gpu_reference"] = {"what": "the reference's compiled contour_cuda get_fields_intersection_inter_slice (oracle/_ref, "
                                            "intersection.cu:227-287) on the same B200, same inputs",
                                    "ms_per_call": ms_ref, "speedup_of_ours": ms_ref / ms_single,
                                    "mask_identical": bool(torch.equal(rm, m0)), "indices_identical": bool(torch.equal(ri, i0))}
    return out


def bench_hessian_sweep(net, dev, peaks, rank, world):
    """BASELINE.json configs[4]: synthetic 16M-point sweep with Hessian for the min-curvature field (SURVEY.md 8d cfg 5):
    value + gradient + Hessian from the forward-over-reverse sweep of the CTA-pair engine, then the closed-form principal curvatures.
    N > 1: points sharded, principal curvatures and directions all-gathered (strong scaling)."""
    import torch
    import inf3dp
    from inf3dp.dist import shard_range, sharded_curvature_sweep
    n = 16_777_216
    g = torch.Generator(device="cpu").manual_seed(0)
    chunk = 1 << 22
    lo, hi, _ = shard_range(n, rank, world)
    x = torch.empty((hi - lo, 3), dtype=torch.float32, device=dev)
    for h in range(0, n, chunk):   # ~U([-1,1]^3), generated on the host generator of SURVEY.md cfg 5 in pieces
        piece = torch.rand((min(chunk, n - h), 3), generator=g) * 2 - 1
        a, b = max(h, lo), min(h + piece.shape[0], hi)
        if a < b:
            x[a - lo:b - lo] = piece[a - h:b - h].to(dev)

    sweep = lambda: sharded_curvature_sweep(net, x, presharded=True, n_total=n)   # order-2 query + curvature tail (+ all-gather)
    ms, (kap, dirs) = _timeit(sweep, dev, 2, 3)
    ms = _max_over_ranks(ms, dev, world)
    flop = 3_938_816        # Hessian mode as a forward jet of 1 + 3 + 6 columns (the figure the survey states)
    flop_for = 4 * 790_528  # what runs: forward-over-reverse, (1 + 3 tangent) rows x the reverse-mode value+gradient sweep
    tf, tf_for = n * flop / (ms / 1e3) / 1e12, n * flop_for / (ms / 1e3) / 1e12
    out = {"metric": "siren_value_grad_hessian_curvature_points_per_s", "value": n / (ms / 1e3), "unit": "points/s",
           "ms_per_sweep": ms, "points": n,
           "kernel": "siren_rev_kernel<CTAS=2,KIND=3> (forward-over-reverse: primal + 3 tangent rows through the reverse-mode "
                     "sweep) + principal_curvatures_kernel",
           "flop_per_point": flop, "achieved_tflops": tf, "frac_of_sustained_peak": tf / peaks["tf_sustained"] / world,
           "flop_per_point_forward_over_reverse": flop_for, "achieved_tflops_forward_over_reverse": tf_for,
           "frac_of_sustained_peak_forward_over_reverse": tf_for / peaks["tf_sustained"] / world,
           "checksum": float(kap[:: n // 4096].double().sum())}
    if world > 1:
        out.update({"n_gpus": world, "scaling": f"strong: 16 777 216 points sharded over {world} ranks, NCCL all-gather of the "
                                                 "curvatures and directions inside the timed region"})
    return out


def bench_collision(sdf_net, dev, rank, world, n_waypoints=4_194_304, m_nozzle=64):
    """BASELINE.json configs[3]: 4M-waypoint TV-SDF + quaternion-field collision query (SURVEY.md 8d cfg 4, M = 64 nozzle
    points per waypoint): quaternion field -> posed nozzle clouds -> SDF / slice / level queries -> predicates / Eq. 27 /
    two-step push (level_collision.py:73-237), device resident.  N > 1: the waypoints are sharded (level_collision.py:259-281 is
    per waypoint), per-waypoint depth and mask all-gathered (strong scaling).
    Synthetic fixtures: waypoints uniform on the torus surface (R 0.6, r 0.3), levels from 8 z-bands, nozzle = 64 points
    of a 1/150-scaled cone shell, networks random-init (seed 2048 order); the SDF output bias is shifted by -0.06 so
    that about half of the cube is inside the part (the raw random-init SDF is positive everywhere)."""
    import copy
    import math
    import torch
    import inf3dp
    from inf3dp.dist import shard_range, sharded_quat_collision_sweep
    torch.manual_seed(2048)
    nets = [inf3dp.SingleBVPNet(type="sine", in_features=3, out_features=o) for o in (1, 1, 1, 1, 4)]
    slice_net, quat_net = nets[1].to(dev), nets[4].to(dev)
    level = inf3dp.MLPClassifier(num_classes=8).to(dev)
    sdf = copy.deepcopy(sdf_net)
    with torch.no_grad():
        sdf.net.net[4][0].bias -= 0.06
    g = torch.Generator(device="cpu").manual_seed(4)
    u = torch.rand(n_waypoints, generator=g) * 2 * math.pi
    v = torch.rand(n_waypoints, generator=g) * 2 * math.pi
    way_all = torch.stack([(0.6 + 0.3 * torch.cos(v)) * torch.cos(u), (0.6 + 0.3 * torch.cos(v)) * torch.sin(u), 0.3 * torch.sin(v)], 1)
    maxsl_all = torch.rand((n_waypoints, 8), generator=g) * 0.1 - 0.05
    lo, hi, _ = shard_range(n_waypoints, rank, world)
    way, maxsl = way_all[lo:hi].to(dev), maxsl_all[lo:hi].to(dev)
    del way_all, maxsl_all
    levels = ((way[:, 2] + 0.3) / 0.6 * 8).clamp(0, 7).long()
    t = torch.linspace(0, 1, m_nozzle)
    nozzle = (torch.stack([0.5 * t * torch.cos(t * 40), 0.5 * t * torch.sin(t * 40), 0.2 + 4.0 * t], 1) / 150.0 * 20.0).to(dev)

    sweep = lambda: sharded_quat_collision_sweep(way, levels, maxsl, nozzle, quat_net, sdf, slice_net, level, 8, -0.9,
                                                 presharded=True, n_total=n_waypoints)
    ms, (depth, hit) = _timeit(sweep, dev, 1, 1)
    ms = _max_over_ranks(ms, dev, world)
    out = {"metric": "collision_waypoints_per_s", "value": n_waypoints / (ms / 1e3), "unit": "waypoints/s", "ms_per_sweep": ms,
           "waypoints": n_waypoints, "nozzle_points_per_waypoint": m_nozzle,
           "nozzle_points_per_s": n_waypoints * m_nozzle / (ms / 1e3), "colliding_waypoints": int(hit.sum()),
           "depth_checksum": float(depth[:: max(1, n_waypoints // 4096)].double().sum()),
           "api": "inf3dp.quat_collision_sweep (quaternion field + level_collision.CollTVSDF forward, device resident)"}
    if world > 1:
        out.update({"n_gpus": world, "scaling": f"strong: {n_waypoints} waypoints sharded over {world} ranks, NCCL all-gather of the "
                                                 "per-waypoint depth and mask inside the timed region"})
    # SURVEY.md 8d cfg 4 at its stated nozzle: data/rect_nozzle_shell.xyz[::10] scaled by 1/150 (level_collision.py:301,463,
    # configs/collision_config.yaml:118) -- M = 6 419 points per waypoint -- on a 65 536-waypoint slice (420 M nozzle points)
    try:
        import numpy as np
        shell = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "cfg1_golden.npz"))["nozzle_shell_every10_scaled"]).to(dev)
        nw = 65536
        a, b, _ = shard_range(nw, rank, world)

        sweep_shell = lambda: sharded_quat_collision_sweep(way[a:b], levels[a:b], maxsl[a:b], shell, quat_net, sdf, slice_net, level,
                                                           8, -0.9, presharded=True, n_total=nw, waypoint_chunk=4096)
        ms2, (d2, h2) = _timeit(sweep_shell, dev, 0, 1)
        ms2 = _max_over_ranks(ms2, dev, world)
        out["reference_nozzle"] = {"nozzle": "data/rect_nozzle_shell.xyz[::10] / 150 (6419 points, the reference's test nozzle)",
                                   "waypoints": nw, "nozzle_points_per_waypoint": int(shell.shape[0]), "ms_per_sweep": ms2,
                                   "waypoints_per_s": nw / (ms2 / 1e3), "nozzle_points_per_s": nw * int(shell.shape[0]) / (ms2 / 1e3),
                                   "colliding_waypoints": int(h2.sum())}
    except Exception as e:   # the fixture is committed; never let this optional leg take the headline down
        out["reference_nozzle"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def bench_volume(net, dev, peaks):
    """fused grid sweep (no coordinate tensor), marching cubes on the resulting 512^3 volume, and the
    three-field intersection fed by 256^3 sample grids (no corner expansion)."""
    import torch
    import inf3dp
    N = GRID_N
    ms_sweep, sdf = _timeit(lambda: inf3dp.grid_field(net, N, chunk=1 << 23), dev, 1, 1)
    level = float(sdf.flatten()[:: 4099].median())     # the random-init SDF never crosses 0: take a level that exists
    h = 2.0 / (N - 1)
    ms_mc, (v, f, nrm, val) = _timeit(lambda: inf3dp.marching_cubes(sdf, level, (h, h, h)), dev, 1, 3)
    mc_bytes = 4 * N ** 3 + v.numel() * 4 + f.numel() * 4 + nrm.numel() * 4 + val.numel() * 4
    nv, nf = int(v.shape[0]), int(f.shape[0])
    del sdf, v, f, nrm, val
    G = 256
    g = torch.stack(torch.meshgrid(*[torch.linspace(-1, 1, G, device=dev)] * 3, indexing="ij"), -1)
    grids = [(g[..., 2] + 0.1 * torch.sin(3 * g[..., 0])).contiguous(), (g[..., 0] + 0.2 * g[..., 1] ** 2).contiguous(),
             (g[..., 1] - 0.15 * g[..., 2] * g[..., 0]).contiguous()]
    del g
    Ns, N1, N2 = 500, 128, 128
    isos = [torch.linspace(-0.9, 0.9, Ns, device=dev), torch.linspace(-0.8, 0.8, N1, device=dev),
            torch.linspace(-0.8, 0.8, N2, device=dev)]
    ms_is, out = _timeit(lambda: inf3dp.get_fields_intersection_from_grids(*grids, *isos, sync=False), dev, 1, 3)
    cells = Ns * N1 * N2
    is_bytes = 12 * G ** 3 + 4 * (Ns + N1 + N2) + 19 * cells
    return {
        "grid_sweep": {"metric": "siren_value_points_per_s", "value": N ** 3 / (ms_sweep / 1e3), "unit": "points/s",
                       "ms": ms_sweep, "grid": N, "api": "inf3dp.grid_field (device query-grid generator + fused value launches, "
                       "no [N^3,3] coordinate tensor)"},
        "marching_cubes": {"metric": "marching_cubes_ms", "ms_per_call": ms_mc, "grid": N, "level": level, "vertices": nv,
                           "triangles": nf, "launches_per_call": 4,
                           "roofline": {"bound": "hbm", "achieved": mc_bytes / (ms_mc / 1e3) / 1e9, "peak": peaks["hbm"],
                                        "unit": "GB/s", "frac": mc_bytes / (ms_mc / 1e3) / 1e9 / peaks["hbm"],
                                        "bytes_algorithmic": mc_bytes, "traffic": None,
                                        "note": "count kernel is instruction bound (profiles/r1c_volume_ncu_summary.md)"}},
        "intersection_on_grids": {"metric": "intersection_ms", "ms_per_call": ms_is, "grid": G, "isovalues": [Ns, N1, N2],
                                  "cells_hit": int(out[0].sum()), "launches_per_call": 4,
                                  "roofline": {"bound": "hbm", "achieved": is_bytes / (ms_is / 1e3) / 1e9, "peak": peaks["hbm"],
                                               "unit": "GB/s", "frac": is_bytes / (ms_is / 1e3) / 1e9 / peaks["hbm"],
                                               "bytes_algorithmic": is_bytes, "traffic": None,
                                               "note": "grid mode reads 12 B per voxel instead of 108: the same kernels are latency / "
                                                       "instruction bound here, not HBM bound; replaces 3 x 8x corner expansion (1.6 GB) "
                                                       "+ centre tensor + ext call"}},
    }


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import inf3dp
    from inf3dp.query import gen_voxel_queries, HostPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the inf3dp hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    engine = os.environ.get("INF3DP_ENGINE", "tc_reverse")  # parity-grade reverse-mode engine is the headline
    kernel_name, flop_per_point, issued_flop_per_point, dtype_note = ENGINES[engine]
    warmup = max(args.warmup, 3)

    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    net.load_state_dict({f"net.net.{i}.0.{k}": torch.from_numpy(np.asarray(v)) for i, (W, b) in enumerate(golden_layers())
                         for k, v in (("weight", W), ("bias", b))})
    net.to(dev)

    # every rank owns one 512^3 grid of the N * 512^3 point job (weak scaling); coordinates resident in HBM
    coords, _, _ = gen_voxel_queries(GRID_N, device=dev)
    n = coords.shape[0]
    normals = torch.empty((n, 3), dtype=torch.float32, device=dev)
    launches_per_step = 2
    pipe = fg = None
    gather_mode = "none (single GPU)"
    if world > 1:
        want = os.environ.get("INF3DP_BENCH_GATHER", "multicast")     # multicast | p2p | nccl
        ok = torch.zeros(1, dtype=torch.int32, device=dev)
        if want in ("multicast", "p2p") and engine == "tc_reverse":
            try:
                from inf3dp.dist import FusedRowGather
                fg = FusedRowGather(n, dev, prefer=want)
                ok += 1
            except Exception as e:   # symmetric memory unavailable on this box: every rank falls back together
                sys.stderr.write(f"[bench] rank {rank}: fused gather unavailable ({type(e).__name__}: {e}); NCCL all-gather instead\n")
                fg = None
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            fg = None
        if fg is not None:
            modes = [None] * world
            dist.all_gather_object(modes, fg.mode)
            if len(set(modes)) != 1:       # mixed capabilities: everybody uses peer stores
                fg.mc, fg.mode = 0, "p2p"
            gather_mode = ("fused into the query kernel's epilogue: " +
                           ("one multimem.st per row to the NVLink multicast address" if fg.mode == "multicast"
                            else f"one 16-byte peer store per row and rank ({world} ranks)") + ", symmetric memory, no SM reserved")
            launches_per_step = 2 + 2     # query + normals + two barrier kernels
        else:
            # NCCL path: the rank's grid is evaluated in pieces and piece c is all-gathered on a side stream while piece c + 1
            # is computed; 4 SMs are left to the NCCL kernels
            from inf3dp.dist import PipelinedAllGather, reserve_sms
            reserve_sms(int(os.environ.get("INF3DP_BENCH_RESERVE_SMS", "4")))
            pipe = PipelinedAllGather(n, 4, chunks=int(os.environ.get("INF3DP_BENCH_CHUNKS", "16")), device=dev)
            launches_per_step = 2 * pipe.chunks
            gather_mode = f"NCCL all_gather_into_tensor in {pipe.chunks} pieces overlapped with the next piece's query (4 SMs left to NCCL)"

    def step():
        if world == 1:
            y, J, _ = net.query(coords, order=1, mode=engine)       # launch 1: fused value + gradient
            inf3dp.unit_normals(J.view(n, 3), out=normals)           # launch 2: normals
            return
        if fg is not None:
            fg.begin()
            fg.query(net, coords)                                    # value + gradient rows stored into every rank's buffer
            inf3dp.unit_normals_rows(fg.local_rows(), out=normals)
            fg.finish()
            return
        for c, (lo, hi) in enumerate(pipe.ranges):
            if engine == "tc_reverse":   # (value, grad) rows written straight into the all-gather's staging buffer
                rows = net.query_rows(coords[lo:hi], out=pipe.slot(c))
                inf3dp.unit_normals_rows(rows, out=normals[lo:hi])
                pipe.gather(c)
            else:
                y, J, _ = net.query(coords[lo:hi], order=1, mode=engine)
                inf3dp.unit_normals(J.view(hi - lo, 3), out=normals[lo:hi])
                pipe.submit(c, y, J.view(hi - lo, 3))
        pipe.finish()

    # kernel-only timing of the dominant kernel (for the roofline), same stream
    def jet_only():
        return net.query(coords, order=1, mode=engine)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        # every rank evaluates the same grid with the same weights: the slab gathered from the neighbour must equal ours
        src = fg if fg is not None else pipe
        mine, theirs = src.rows_of_rank(rank), src.rows_of_rank((rank + 1) % world)
        pick = torch.arange(0, n, max(1, n // 65536), device=dev)
        if not torch.equal(mine[pick], theirs[pick]):
            raise SystemExit(f"bench.py: rank {rank}: all-gathered slab differs from the local one")
        want_rows = net.query_rows(coords[pick].contiguous())
        if not torch.equal(mine[pick], want_rows):
            raise SystemExit(f"bench.py: rank {rank}: gathered rows differ from a plain query of the same points")
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms_step = e0.elapsed_time(e1) / args.steps

    # dominant kernel alone (CUDA events on the launching stream)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        jet_only()
    k1.record()
    torch.cuda.synchronize(dev)
    ms_kernel = k0.elapsed_time(k1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    if pipe is not None:
        from inf3dp.dist import reserve_sms
        reserve_sms(0)

    # ---- end to end through the public host-buffer API -------------------------------------------------------
    del normals
    pipe = None
    coords_host = torch.empty((n, 3), dtype=torch.float32, pin_memory=True)
    coords_host.copy_(coords)
    out_host = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    if world > 1 and fg is not None:
        from inf3dp.dist import ShardedHostPipeline
        del fg
        hp = ShardedHostPipeline(net, n, chunk=1 << 23, device=dev, prefer="multicast" if "multimem" in gather_mode else "p2p")
        src_rank = (rank + 1) % world
        run_e2e = lambda: hp.run(coords_host, out_host, source_rank=src_rank)
        e2e_api = ("inf3dp.dist.ShardedHostPipeline.run(pinned coords of this rank's shard, pinned out): H2D -> fused query + "
                   "all-gather over NVLink -> D2H of the NEIGHBOUR rank's rows (the collective is on the timed path)")
    else:
        hp = HostPipeline(net, chunk=1 << 23, device=dev, mode=engine)
        run_e2e = lambda: hp.run(coords_host, out_host)
        e2e_api = "inf3dp.HostPipeline.run(pinned coords, pinned out)" + (" per rank, no collective (NCCL fallback path)" if world > 1 else "")
    for _ in range(3):
        run_e2e()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        run_e2e()   # returns after the last D2H has landed
    g1.record()
    torch.cuda.synchronize(dev)
    ms_e2e = max(g0.elapsed_time(g1), (time.perf_counter() - t0) * 1e3) / args.steps
    # the host result must be the device-resident result: compare a strided sample (every rank evaluates the same grid, so
    # the neighbour's rows equal ours) and abort on a mismatch
    stride = max(1, n // 262144)
    want_rows = net.query_rows(coords[::stride].contiguous()).cpu()
    if not torch.equal(out_host[::stride], want_rows):
        bad = int((out_host[::stride] != want_rows).any(dim=1).sum())
        raise SystemExit(f"bench.py: rank {rank}: e2e host result differs from the device-resident query on {bad} of {want_rows.shape[0]} sampled rows")
    e2e_check = {"rows_compared": int(want_rows.shape[0]), "bit_identical": True}
    del hp, coords_host, out_host, want_rows

    # max over ranks
    rank_ms = None
    if world > 1:
        mine_ms = torch.tensor([ms_step], dtype=torch.float64, device=dev)
        all_ms = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(all_ms, mine_ms)
        rank_ms = [round(float(v), 3) for v in all_ms.tolist()]
        tt = torch.tensor([ms_step, ms_kernel, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_step, ms_kernel, ms_e2e = [float(v) for v in tt.tolist()]
    del coords
    torch.cuda.empty_cache()

    # The other configurations of BASELINE.json.  N = 1: plain.  N > 1: sharded the way SURVEY.md 8e says, with the
    # all-gather of their results inside the timed region (strong scaling), each checked against the single-GPU result.
    extras = (world == 1 and rank == 0)
    iso = bench_isocontours(dev, peaks, rank, world, extras)
    torch.cuda.empty_cache()
    isect = bench_intersection(dev, peaks, rank, world, extras)
    torch.cuda.empty_cache()
    hess = bench_hessian_sweep(net, dev, peaks, rank, world)
    torch.cuda.empty_cache()
    coll = bench_collision(net, dev, rank, world)
    torch.cuda.empty_cache()

    if rank == 0:
        total_pts = world * n
        value = total_pts / (ms_step / 1e3)
        kernel_tflops = n * flop_per_point / (ms_kernel / 1e3) / 1e12
        # the jet kernel runs ~0.5-1 s per launch under the power cap: the sustained cuBLAS figure is the fair peak
        peak = peaks["tf_sustained"]
        vol = bench_volume(net, dev, peaks) if world == 1 else None
        cores = os.cpu_count() or 1
        cpu_base = None
        if world == 1:   # reported baseline: rank 0 at N=1 only (torchrun pins OMP threads to 1 for N>1)
            cpu_pps, cpu_dt = cpu_reference_points_per_s(CPU_SAMPLE, 2, cores)
            cpu_base = {"value": cpu_pps, "unit": "points/s", "cores": cores, "kind": _CPU_REF["kind"],
                        "sample": f"{CPU_SAMPLE} consecutive points from the middle of the 512^3 grid, best of 2, {_CPU_REF['what']}, "
                                  "max_batch 32^3 (torch CPU)"}
        line = {
            "metric": "siren_value_grad_points_per_s", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "ms_per_step_by_rank": rank_ms,
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype_note,
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "network": "3-256-256-256-256-1 sine, random init seed 2048",
                       "grid": GRID_N, "points_per_gpu": n, "engine": engine,
                       "l2_note": "inputs (1.6 GB coords) and outputs (2.1 GB) per step exceed the 126 MB L2",
                       "parallelism": (f"dp{world}: points sharded, weights replicated, all-gather of the (value, grad) rows "
                                       f"{gather_mode}") if world > 1 else "single GPU"},
            "e2e": {"value": total_pts / (ms_e2e / 1e3), "unit": "points/s", "h2d_bytes_per_step": n * 12,
                    "d2h_bytes_per_step": n * 16, "ms_per_step": ms_e2e, "api": e2e_api, "check_vs_device_resident": e2e_check},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "tensor", "achieved": kernel_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": kernel_tflops / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one bench-sized launch from a committed ncu capture of
                         # THIS source (profiles/traffic.json), else null; algorithmic HBM bytes are 28 per point = 3.76e9
                         "traffic": measured_traffic("siren_rev", "siren_rev.cu") if engine == "tc_reverse" else None,
                         "traffic_unit": "DRAM bytes per launch (ncu, profiles/traffic.json)",
                         "kernel": kernel_name,
                         "kernel_ms": ms_kernel, "flop_per_point": flop_per_point,
                         "formulation": "reverse-mode (forward + backward sweep)" if engine == "tc_reverse" else "forward-mode jet",
                         "peak_kind": "bf16_tflops_sustained",
                         "peak_source": peaks["source"], "frac_of_burst_peak": kernel_tflops / peaks["tf_burst"],
                         "executed_tensor_flop_per_point": issued_flop_per_point,
                         "tensor_pipe_frac_of_sustained_peak": n * issued_flop_per_point / (ms_kernel / 1e3) / 1e12 / peak,
                         "forward_mode_flop_per_point": FLOP_FORWARD_MODE,
                         "reverse_mode_flop_per_point": FLOP_REVERSE_MODE},
            "cpu_baseline": cpu_base,
            "isocontour": iso,
            "intersection": isect,
            "hessian_sweep": hess,
            "collision_sweep": coll,
            "volume": vol,
            "clocks": clocks,
            "torch": torch.__version__,
        }
        emit_json_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit_json_line(line):
    """The one JSON line goes to the real stdout; everything else the process prints (NCCL banners, library chatter)
    was redirected to stderr at start-up so that stdout carries exactly one line."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)   # C-level stdout (e.g. "NCCL version ...") -> stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
