#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native INF-3DP hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

Metric (BASELINE.json): SIREN value+grad points/s.  Workload at every N: BASELINE.json configs[1], the dense
512^3 grid evaluation with normals of the SDF network (sdf_meshing path): coordinates = the reference's
gen_voxle_queries(512) (134 217 728 points, shear quirk included), network = sdf_config arch
(3->256->256->256->256->1, sine, random init seed 2048), outputs = field value, gradient and unit normal per point.
A step is one pass over the whole grid.  With N ranks every rank evaluates its own 512^3 grid shard of an
N*512^3-point job (weak scaling) and the step ends with one NCCL all-gather of the (value, gradient) rows.

`value`  : whole-job points/s with coordinates resident in HBM (CUDA events, max over ranks).
`e2e`    : the same metric through the public host-buffer API (inf3dp.HostPipeline): pinned host coordinates
           -> H2D -> fused kernel -> D2H of (value, gradient) into pinned host memory, copies inside the timed region.
`roofline`: dominant kernel (siren_rev_kernel, the reverse-mode tcgen05 engine): algorithmic FLOPs of the formulation it
           executes (790 528 per point: one forward + one backward sweep, SURVEY.md 8d "reverse-mode equivalent")
           / launch duration against the measured dense tensor peak in MEASURED_PEAKS.json.  INF3DP_ENGINE=tc_precise
           selects the forward-mode jet engine instead (1 576 448 FLOP per point, the figure it is then judged on).
`cpu_baseline`: the torch-CPU port of the reference's own path (oracle/siren_torch_port.py) on the host cores.
`--impl reference`: that CPU path alone, same metric, bounded sample per step.
Second metric (isocontour layers/s) is carried in the `isocontour` object of the same line.
"""
import argparse
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


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm": float(d["hbm_gbs"]), "tf_burst": float(d["bf16_tflops"]),
                "tf_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


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
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_points_per_s(sample_points, repeats, threads):
    """The reference's torch-CPU path (oracle port): field_query_with_grads on `sample_points` grid points."""
    import numpy as np
    import torch
    from oracle import siren_torch_port as port
    from oracle import siren_oracle as so
    torch.set_num_threads(threads)
    model = port.PortSiren([(W.astype(np.float32), b.astype(np.float32)) for W, b in golden_layers()])
    model.eval()
    # a contiguous slab of the 512^3 grid (same generator as the GPU arm), bounded so the run stays short
    side = 64
    grid = so.gen_voxel_queries(side)
    coords = torch.from_numpy(grid[:sample_points].copy())
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        port.field_query_with_grads(model, coords, max_batch=32 ** 3)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample_points / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 262144
    times = []
    import torch
    for i in range(args.warmup + args.steps):
        pps, dt = cpu_port_points_per_s(sample, 1, cores)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms / 1e3)
    desc = f"{sample} grid points per step (utils.field_query_with_grads, max_batch 32^3) on {cores} host threads"
    line = {"impl": "reference", "metric": "siren_value_grad_points_per_s", "value": value, "unit": "points/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "network": "3-256-256-256-256-1 sine, seed 2048", "grid": GRID_N},
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "torch": torch.__version__}
    emit_json_line(line)


def bench_isocontours(dev, peaks, cpu_baseline=False):
    """BASELINE.json configs[2]: 200-isovalue isocontour extraction on a closed synthetic mesh (SURVEY.md 8d cfg 3)."""
    import numpy as np
    import torch
    import inf3dp
    from oracle import contour_oracle as co
    V, F = co.torus_mesh(1000, 500)           # 1.0 M faces, 0.5 M vertices
    f = (V[:, 2] + 0.25 * np.sin(3.0 * V[:, 0]) * np.cos(2.0 * V[:, 1])).astype(np.float32)
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

    def call():
        inf3dp._lib.check(L.inf3dp_extract_isocontours(*[inf3dp._lib.ptr(a) for a in t], nV, nF, K,
                                                       inf3dp._lib.ptr(merged), inf3dp._lib.ptr(nums),
                                                       inf3dp._lib.ptr(ws), wsb, st))
    for _ in range(3):
        call()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    seg = (inf3dp._lib.ctypes.c_int32 * K)()
    inf3dp._lib.check(L.inf3dp_extract_isocontours_status(inf3dp._lib.ptr(ws), K, seg, st))
    S = int(sum(seg)); C = int(nums.sum().item())
    contract = 12 * nF + 16 * nV + 4 * K + 12 * K * nF + 4 * K
    useful = 12 * nF + 16 * nV + 4 * K + 12 * (S + C) + 4 * K
    cpu = None
    if cpu_baseline:
        # reported baseline (rank 0, N = 1): the single-thread C restatement of the reference kernels on a bounded sample of
        # the same workload (the reference has no CPU implementation of this op); checked against the GPU rows while at it
        sel = np.ascontiguousarray(iso[::25][:8])
        t0 = time.perf_counter()
        mo, no, so_ = co.extract_isocontours(V, F, f, sel)
        dt = time.perf_counter() - t0
        rows = differ = 0
        worst = 0.0
        for j, k in enumerate(range(0, K, 25)[:8]):     # same piece count => same row order: compare row by row
            if int(nums[k]) == int(no[j]):
                mine = merged[k].cpu().numpy()
                used = int(so_[j]) + int(no[j])
                d = np.abs(mine[:used] - mo[j][:used]).max(axis=1) if used else np.zeros(0)
                rows += used
                differ += int((d > 0).sum())
                worst = max(worst, float(d.max()) if used else 0.0)
        cpu = {"value": len(sel) / dt, "unit": "layers/s", "cores": 1, "kind": "port",
               "sample": f"{len(sel)} of the {K} isovalues, oracle/contour_oracle.c (gcc -O2, one thread)",
               # segments shorter than the reference's 1e-6 merge radius can be entered through either end (DESIGN.md divergence 1)
               "check_vs_gpu": {"rows_compared": rows, "rows_not_bit_identical": differ, "max_abs_vertex_diff": worst}}
    return {"metric": "isocontour_layers_per_s", "value": K / (ms / 1e3), "unit": "layers/s", "ms_per_call": ms,
            "cpu_baseline": cpu,
            "config": {"workload": "200-isovalue isocontour extraction, closed torus mesh", "faces": nF,
                       "vertices": nV, "isovalues": K, "segments": S, "loops": C},
            "roofline": {"bound": "hbm", "achieved": contract / (ms / 1e3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                         "frac": contract / (ms / 1e3) / 1e9 / peaks["hbm"],
                         # dram bytes of the nine launches of one call (ncu --set full, profiles/r1b_isocontours_ncu_summary.md)
                         "traffic": 2.46e9, "traffic_unit": "bytes per call (ncu)",
                         "bytes_contract": contract, "bytes_useful": useful,
                         "useful_gbs": useful / (ms / 1e3) / 1e9, "peak_source": peaks["source"]},
            "gpu_launches_per_call": 9}


def bench_hessian_sweep(net, dev, peaks):
    """BASELINE.json configs[4]: synthetic 16M-point sweep with Hessian for the min-curvature field (SURVEY.md 8d cfg 5):
    value + gradient + Hessian from the tensor-core second-order jet, then the closed-form principal curvatures."""
    import torch
    import inf3dp
    n = 16_777_216
    g = torch.Generator(device="cpu").manual_seed(0)
    chunk = 1 << 22
    x = torch.empty((n, 3), dtype=torch.float32, device=dev)
    for h in range(0, n, chunk):   # ~U([-1,1]^3), generated on the host generator of SURVEY.md cfg 5 in pieces
        x[h:h + chunk] = (torch.rand((min(chunk, n - h), 3), generator=g) * 2 - 1).to(dev)

    def sweep():
        y, J, H = net.query(x, order=2)
        return inf3dp.principal_curvatures(J[:, 0], H[:, 0])
    for _ in range(2):
        sweep()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 3
    e0.record()
    for _ in range(iters):
        kap, dirs = sweep()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    flop = 3_938_816   # Hessian mode, 1 + 3 + 6 jet columns
    tf = n * flop / (ms / 1e3) / 1e12
    return {"metric": "siren_value_grad_hessian_curvature_points_per_s", "value": n / (ms / 1e3), "unit": "points/s",
            "ms_per_sweep": ms, "points": n, "kernel": "siren_rev_kernel<CTAS=2,KIND=1> (forward-mode second-order jet) + principal_curvatures_kernel",
            "flop_per_point": flop, "achieved_tflops": tf, "frac_of_sustained_peak": tf / peaks["tf_sustained"],
            "checksum": float(kap[:: n // 4096].double().sum())}


def bench_collision(sdf_net, dev, n_waypoints=4_194_304, m_nozzle=64):
    """BASELINE.json configs[3]: 4M-waypoint TV-SDF + quaternion-field collision query (SURVEY.md 8d cfg 4, M = 64 nozzle
    points per waypoint): quaternion field -> posed nozzle clouds -> SDF value+gradient -> slice value+gradient on the
    inside points -> level MLP -> predicates / Eq. 27 / two-step push (level_collision.py:73-237), device resident.
    Synthetic fixtures: waypoints uniform on the torus surface (R 0.6, r 0.3), levels from 8 z-bands, nozzle = 64 points
    of a 1/150-scaled cone shell, networks random-init (seed 2048 order); the SDF output bias is shifted by -0.06 so
    that about half of the cube is inside the part (the raw random-init SDF is positive everywhere)."""
    import copy
    import math
    import torch
    import inf3dp
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
    way = torch.stack([(0.6 + 0.3 * torch.cos(v)) * torch.cos(u), (0.6 + 0.3 * torch.cos(v)) * torch.sin(u), 0.3 * torch.sin(v)], 1).to(dev)
    levels = ((way[:, 2] + 0.3) / 0.6 * 8).clamp(0, 7).long()
    maxsl = (torch.rand((n_waypoints, 8), generator=g) * 0.1 - 0.05).to(dev)
    t = torch.linspace(0, 1, m_nozzle)
    nozzle = (torch.stack([0.5 * t * torch.cos(t * 40), 0.5 * t * torch.sin(t * 40), 0.2 + 4.0 * t], 1) / 150.0 * 20.0).to(dev)

    def sweep():
        return inf3dp.quat_collision_sweep(way, levels, maxsl, nozzle, quat_net, sdf, slice_net, level, 8, -0.9)
    sweep()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    depth, hit = sweep()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    return {"metric": "collision_waypoints_per_s", "value": n_waypoints / (ms / 1e3), "unit": "waypoints/s", "ms_per_sweep": ms,
            "waypoints": n_waypoints, "nozzle_points_per_waypoint": m_nozzle,
            "nozzle_points_per_s": n_waypoints * m_nozzle / (ms / 1e3), "colliding_waypoints": int(hit.sum()),
            "api": "inf3dp.quat_collision_sweep (quaternion field + level_collision.CollTVSDF forward, device resident)"}


def bench_volume(net, dev, peaks):
    """fused grid sweep (no coordinate tensor), marching cubes on the resulting 512^3 volume, and the
    three-field intersection fed by 256^3 sample grids (no corner expansion)."""
    import torch
    import inf3dp

    def timeit(fn, reps):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps, out

    N = GRID_N
    ms_sweep, sdf = timeit(lambda: inf3dp.grid_field(net, N, chunk=1 << 23), 1)
    level = float(sdf.flatten()[:: 4099].median())     # the random-init SDF never crosses 0: take a level that exists
    h = 2.0 / (N - 1)
    ms_mc, (v, f, nrm, val) = timeit(lambda: inf3dp.marching_cubes(sdf, level, (h, h, h)), 3)
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
    ms_is, out = timeit(lambda: inf3dp.get_fields_intersection_from_grids(*grids, *isos, sync=False), 3)
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
                                               "note": "replaces 3 x 8x corner expansion (1.6 GB) + centre tensor + ext call"}},
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

    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    net.load_state_dict({f"net.net.{i}.0.{k}": torch.from_numpy(np.asarray(v)) for i, (W, b) in enumerate(golden_layers())
                         for k, v in (("weight", W), ("bias", b))})
    net.to(dev)

    # every rank owns one 512^3 grid of the N * 512^3 point job (weak scaling); coordinates resident in HBM
    coords, _, _ = gen_voxel_queries(GRID_N, device=dev)
    n = coords.shape[0]
    normals = torch.empty((n, 3), dtype=torch.float32, device=dev)
    launches_per_step = 2
    pipe = None
    if world > 1:
        # the one collective of the sharded path, overlapped: the rank's grid is evaluated in 8 pieces and piece c is
        # all-gathered on a side stream while piece c + 1 is computed; 4 SMs are left to the NCCL kernels
        from inf3dp.dist import PipelinedAllGather, reserve_sms
        reserve_sms(int(os.environ.get("INF3DP_BENCH_RESERVE_SMS", "4")))
        pipe = PipelinedAllGather(n, 4, chunks=int(os.environ.get("INF3DP_BENCH_CHUNKS", "8")), device=dev)
        launches_per_step = 2 * pipe.chunks

    def step():
        if pipe is None:
            y, J, _ = net.query(coords, order=1, mode=engine)       # launch 1: fused value + gradient
            inf3dp.unit_normals(J.view(n, 3), out=normals)           # launch 2: normals
            return y, J
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

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        # every rank evaluates the same grid with the same weights: the slab gathered from the neighbour must equal ours
        mine, theirs = pipe.rows_of_rank(rank), pipe.rows_of_rank((rank + 1) % world)
        pick = torch.arange(0, n, max(1, n // 65536), device=dev)
        if not torch.equal(mine[pick], theirs[pick]):
            raise SystemExit(f"bench.py: rank {rank}: all-gathered slab differs from the local one")
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
    if world > 1:
        from inf3dp.dist import reserve_sms
        reserve_sms(0)

    # ---- end to end through the public host-buffer API -------------------------------------------------------
    del normals
    coords_host = torch.empty((n, 3), dtype=torch.float32, pin_memory=True)
    coords_host.copy_(coords)
    out_host = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    pipe = HostPipeline(net, chunk=1 << 23, device=dev, mode=engine)
    for _ in range(3):
        pipe.run(coords_host, out_host)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        pipe.run(coords_host, out_host)   # returns after the last D2H has landed
    g1.record()
    torch.cuda.synchronize(dev)
    ms_e2e = max(g0.elapsed_time(g1), (time.perf_counter() - t0) * 1e3) / args.steps
    e2e_checksum = float(out_host[:: max(1, n // 4096), 0].double().sum())

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

    # The other configurations of BASELINE.json (isovalues, Hessian sweep, collision waypoints) shard without any exchange:
    # at N > 1 every rank runs its own copy of the workload (weak scaling, rank-local consumers), the time is the maximum
    # over ranks and the reported value is the aggregate of all ranks.
    iso = bench_isocontours(dev, peaks, cpu_baseline=(world == 1 and rank == 0))
    hess = bench_hessian_sweep(net, dev, peaks)
    coll = bench_collision(net, dev)
    if world > 1:
        tt = torch.tensor([iso["ms_per_call"], hess["ms_per_sweep"], coll["ms_per_sweep"]], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        for obj, key, ms_max in ((iso, "ms_per_call", float(tt[0])), (hess, "ms_per_sweep", float(tt[1])),
                                 (coll, "ms_per_sweep", float(tt[2]))):
            scale = world * obj[key] / ms_max          # value was computed from this rank's own time
            obj["value"] *= scale
            obj[key] = ms_max
            obj["n_gpus"] = world
            obj["scaling"] = "weak: one copy of the workload per rank, no exchange, max time over ranks"
        if "nozzle_points_per_s" in coll:
            coll["nozzle_points_per_s"] = coll["value"] * coll["nozzle_points_per_waypoint"]

    if rank == 0:
        total_pts = world * n
        value = total_pts / (ms_step / 1e3)
        kernel_tflops = n * flop_per_point / (ms_kernel / 1e3) / 1e12
        # the jet kernel runs ~0.5-1 s per launch under the power cap: the sustained cuBLAS figure is the fair peak
        peak = peaks["tf_sustained"]
        vol = bench_volume(net, dev, peaks)
        cores = os.cpu_count() or 1
        cpu_base = None
        if world == 1:   # reported baseline: rank 0 at N=1 only (torchrun pins OMP threads to 1 for N>1)
            cpu_pps, cpu_dt = cpu_port_points_per_s(262144, 2, cores)
            cpu_base = {"value": cpu_pps, "unit": "points/s", "cores": cores, "kind": "port",
                        "sample": "262144 grid points, best of 2, oracle/siren_torch_port.field_query_with_grads (torch CPU)"}
        line = {
            "metric": "siren_value_grad_points_per_s", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "ms_per_step_by_rank": rank_ms,
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype_note,
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "network": "3-256-256-256-256-1 sine, random init seed 2048",
                       "grid": GRID_N, "points_per_gpu": n, "engine": engine,
                       "l2_note": "inputs (1.6 GB coords) and outputs (2.1 GB) per step exceed the 126 MB L2",
                       "parallelism": (f"dp{world}: points sharded, weights replicated, one all-gather of the outputs in 8 pieces overlapped "
                                       f"with the next piece's query (4 SMs left to NCCL)") if world > 1 else "single GPU"},
            "e2e": {"value": total_pts / (ms_e2e / 1e3), "unit": "points/s", "h2d_bytes_per_step": n * 12,
                    "d2h_bytes_per_step": n * 16, "ms_per_step": ms_e2e, "api": "inf3dp.HostPipeline.run(pinned coords, pinned out)",
                    "checksum": e2e_checksum},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "tensor", "achieved": kernel_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": kernel_tflops / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one bench-sized launch (ncu --set full,
                         # profiles/r1b_siren_rev_ncu_summary.md); algorithmic HBM bytes are 28 per point = 3.76e9
                         "traffic": 3.825e9 if engine == "tc_reverse" else None, "traffic_unit": "bytes per launch (ncu)",
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
            "hessian_sweep": hess,
            "collision_sweep": coll,
            "volume": vol,
            "clocks": clocks,
            "torch": torch.__version__,
        }
        emit_json_line(line)
    if world > 1:
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
