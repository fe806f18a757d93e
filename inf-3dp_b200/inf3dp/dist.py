"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink), replicated weights, points or
isovalues sharded contiguously, one final all-gather.  The reference has no distributed code.
Works with the gloo backend on CPU tensors for host-side tests of the sharding logic.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous ceil(n/world) ranges; the last ranks may be short or empty."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


def all_gather_rows(local, n_total, group=None):
    """Gather per-rank row slabs (each rank holds rows shard_range(n_total, rank, world)) into [n_total, ...]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    lo, hi, per = shard_range(n_total, rank, world)
    if local.shape[0] != hi - lo:
        raise RuntimeError(f"all_gather_rows: rank {rank} passed {local.shape[0]} rows, its shard of {n_total} is [{lo}, {hi})")
    if local.dtype in (torch.int16, torch.bool, torch.uint16):   # not NCCL dtypes: gather the rows as bytes
        tail = tuple(local.shape[1:])
        width = 1
        for t in tail:
            width *= int(t)
        raw = local.contiguous().view(hi - lo, width).view(torch.uint8)      # explicit width: an empty shard has no -1
        return all_gather_rows(raw, n_total, group).view(local.dtype).view((n_total,) + tail)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:hi - lo] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return out[:n_total]


def sharded_value_grad(net, coords, mode=None, group=None, gather=True):
    """Each rank evaluates its contiguous point range; optionally all-gathers [N,4] = (value, grad)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = coords.shape[0]
    lo, hi, _ = shard_range(n, rank, world)
    y, J, _ = net.query(coords[lo:hi], order=1, mode=mode)
    local = torch.cat([y[:, :1], J[:, 0, :]], dim=1)
    return all_gather_rows(local, n, group) if gather else local


def sharded_isocontours(vertices, faces, fields, iso_values, group=None, gather=True):
    """Isovalues are sharded (chaining needs all segments of one isovalue on one rank); mesh and field replicated."""
    from .contours import extract_isocontours_raw
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    K = iso_values.numel()
    lo, hi, _ = shard_range(K, rank, world)
    merged, nums = extract_isocontours_raw(vertices, faces, fields, iso_values[lo:hi].contiguous())
    if not gather or world == 1:
        return merged, nums
    return all_gather_rows(merged, K, group), all_gather_rows(nums, K, group)


def sharded_isocontours_compact(vertices, faces, fields, iso_values, group=None):
    """Isovalue-sharded extraction with COMPACT exchange: every rank extracts its K / world isovalues as compact row
    streams (inf3dp.extract_isocontours_compact) and only the used rows travel -- tens of MB instead of the 12 K F bytes of
    the dense contract, which is what makes sharding the isovalues pay.
    -> rows [R,3] f32, row_offsets [K] i64, row_counts [K] i32, contour_nums [K] i32, identical on every rank:
    rows[row_offsets[k] : row_offsets[k] + row_counts[k]] is the row stream of isovalue k."""
    from .contours import extract_isocontours_compact
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    K = iso_values.numel()
    lo, hi, per = shard_range(K, rank, world)
    rows, offs, counts, nums = extract_isocontours_compact(vertices, faces, fields, iso_values[lo:hi].contiguous())
    if world == 1:
        return rows, offs, counts, nums
    dev = rows.device
    mine = (offs + counts).max().reshape(1) if hi > lo else torch.zeros(1, dtype=torch.int64, device=dev)
    totals = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(totals, mine.to(torch.int64), group=group)
    totals_h = [int(t) for t in totals.tolist()]          # one small host sync: the gather buffers are sized from it
    cap = max(1, max(totals_h))
    send = torch.zeros((cap, 3), dtype=torch.float32, device=dev)
    send[:totals_h[rank]] = rows[:totals_h[rank]]
    recv = torch.empty((world, cap, 3), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(recv.view(world * cap, 3), send, group=group)
    base = [0]
    for t in totals_h[:-1]:
        base.append(base[-1] + t)
    rows_all = torch.cat([recv[r, :totals_h[r]] for r in range(world)], dim=0)
    counts_all = all_gather_rows(counts, K, group)
    nums_all = all_gather_rows(nums, K, group)
    offs_all = all_gather_rows(offs, K, group)
    shift = torch.zeros(K, dtype=torch.int64, device=dev)
    for r in range(world):
        a, b, _ = shard_range(K, r, world)
        shift[a:b] = base[r]
    return rows_all, offs_all + shift, counts_all, nums_all


def _world_rank(group):
    if not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def sharded_fields_intersection(slice_fields, lattice1_fields, lattice2_fields, slice_iso_values, lattice1_iso_values,
                                lattice2_iso_values, voxel_centers, group=None, gather=True, sync=False):
    """get_fields_intersection_inter_slice with the SLICE isovalues sharded (output dim 0; intersection.cu:227-287 treats
    every (slice, lattice1, lattice2) cell independently).  The field grids are replicated, every rank intersects its
    ceil(Ns / world) slice isovalues with all lattice isovalues, and the three outputs are all-gathered on dim 0.  The
    slice component of `isovalue_indices` is local to the shard as the kernel writes it; it is shifted to the global
    slice index where the cell holds an intersection (>= 0), the -1 of empty cells stays.
    -> [inter_mask bool [Ns,N1,N2], isovalue_indices int16 [Ns,N1,N2,3], intersections f32 [Ns,N1,N2,3]], identical on
    every rank and identical to the single-GPU call (gather=False: the local [hi-lo,N1,N2,...] slabs, indices global)."""
    from . import contours
    world, rank = _world_rank(group)
    Ns, N1, N2 = slice_iso_values.numel(), lattice1_iso_values.numel(), lattice2_iso_values.numel()
    lo, hi, _ = shard_range(Ns, rank, world)
    if hi > lo:
        m, i, p = contours.get_fields_intersection_inter_slice(slice_fields, lattice1_fields, lattice2_fields,
                                                               slice_iso_values[lo:hi].contiguous(), lattice1_iso_values,
                                                               lattice2_iso_values, voxel_centers, sync=sync)
    else:   # more ranks than slice isovalues: this rank contributes no rows
        dev = slice_fields.device
        m = torch.zeros((0, N1, N2), dtype=torch.bool, device=dev)
        i = torch.zeros((0, N1, N2, 3), dtype=torch.int16, device=dev)
        p = torch.zeros((0, N1, N2, 3), dtype=torch.float32, device=dev)
    if world == 1:
        return [m, i, p]
    i = i.clone()
    i[..., 0] += lo * (i[..., 0] >= 0).to(i.dtype)
    if not gather:
        return [m, i, p]
    return [all_gather_rows(m.view(hi - lo, N1 * N2), Ns, group).view(Ns, N1, N2),
            all_gather_rows(i.view(hi - lo, N1 * N2 * 3), Ns, group).view(Ns, N1, N2, 3),
            all_gather_rows(p.view(hi - lo, N1 * N2 * 3), Ns, group).view(Ns, N1, N2, 3)]


def sharded_quat_collision_sweep(waypts, waypt_levels, current_level_maxslice, nozzle_pcd, quat_decoder, sdf_decoder,
                                 slice_decoder, level_decoder, num_classes, base_th, group=None, gather=True, presharded=False,
                                 n_total=None, **kw):
    """inf3dp.quat_collision_sweep with the WAYPOINTS sharded (level_collision.py:259-281 evaluates every waypoint on its
    own: quaternion field -> posed nozzle cloud -> TV-SDF predicates -> per-waypoint depth and mask).  The five networks
    and the nozzle cloud are replicated; every rank sweeps waypoints shard_range(N, rank, world) and the per-waypoint
    depth [N] and mask [N] are all-gathered.
    presharded=False: `waypts`, `waypt_levels`, `current_level_maxslice` hold all N waypoints on every rank (this rank's
    rows are sliced out here).  presharded=True: they hold this rank's shard only and `n_total` = N.
    -> (depth f32 [N], hit bool [N]), identical on every rank (gather=False: the local shard's)."""
    from . import collision
    world, rank = _world_rank(group)
    if presharded:
        if n_total is None:
            raise RuntimeError("sharded_quat_collision_sweep: presharded=True needs n_total")
        n = int(n_total)
        lo, hi, _ = shard_range(n, rank, world)
        if waypts.shape[0] != hi - lo:
            raise RuntimeError(f"sharded_quat_collision_sweep: rank {rank} holds {waypts.shape[0]} waypoints, its shard of "
                               f"{n} is [{lo}, {hi})")
        w, l, s = waypts, waypt_levels, current_level_maxslice
    else:
        n = waypts.shape[0]
        lo, hi, _ = shard_range(n, rank, world)
        w, l, s = waypts[lo:hi], waypt_levels[lo:hi], current_level_maxslice[lo:hi]
    d, h = collision.quat_collision_sweep(w, l, s, nozzle_pcd, quat_decoder, sdf_decoder, slice_decoder, level_decoder,
                                          num_classes, base_th, **kw)
    if world == 1 or not gather:
        return d, h
    return all_gather_rows(d, n, group), all_gather_rows(h, n, group)


def sharded_curvature_sweep(net, coords, group=None, gather=True, presharded=False, n_total=None, mode=None):
    """The min / max-curvature field of a SIREN over N points with the POINTS sharded (SURVEY.md 8d cfg 5): value +
    gradient + Hessian of this rank's points (order-2 query), the closed-form principal curvatures and directions
    (diff_operators.py:27-44 + utils.py min_max_curvs_and_vectors), then one all-gather of [N,2] and [N,2,3].
    presharded as in sharded_quat_collision_sweep.  -> (kappa [N,2], dirs [N,2,3])."""
    from . import diffops
    world, rank = _world_rank(group)
    if presharded:
        if n_total is None:
            raise RuntimeError("sharded_curvature_sweep: presharded=True needs n_total")
        n = int(n_total)
        lo, hi, _ = shard_range(n, rank, world)
        if coords.shape[0] != hi - lo:
            raise RuntimeError(f"sharded_curvature_sweep: rank {rank} holds {coords.shape[0]} points, its shard of {n} is "
                               f"[{lo}, {hi})")
        x = coords
    else:
        n = coords.shape[0]
        lo, hi, _ = shard_range(n, rank, world)
        x = coords[lo:hi]
    if hi > lo:
        _, J, H = net.query(x, order=2, mode=mode)
        kap, dirs = diffops.principal_curvatures(J[:, 0], H[:, 0])
    else:   # more ranks than points
        kap = torch.zeros((0, 2), dtype=torch.float32, device=coords.device)
        dirs = torch.zeros((0, 2, 3), dtype=torch.float32, device=coords.device)
    if world == 1 or not gather:
        return kap, dirs
    return all_gather_rows(kap, n, group), all_gather_rows(dirs.reshape(hi - lo, 6), n, group).view(n, 2, 3)


class PipelinedAllGather:
    """All-gather of per-rank row slabs that overlaps the producer: the local rows are produced in `chunks` pieces, and
    piece c is gathered (on a side stream for CUDA tensors) while piece c + 1 is being computed -- the one collective of
    the sharded path costs the time of its last piece only.

    Every rank holds `n_local` rows of `width` float32 columns.  The result lives in `self.gathered`
    [chunks, world, m, width] (m = ceil(n_local / chunks); the tail of the last piece is padding);
    `rows_of_rank(r)` is the [n_local, width] slab of rank r.

        pipe = PipelinedAllGather(n_local, 4, chunks=8, device=dev)
        for c, (lo, hi) in enumerate(pipe.ranges):
            y, J, _ = net.query(coords[lo:hi], order=1)
            pipe.submit(c, y, J.view(-1, 3))          # columns are concatenated into the staging rows of piece c
        pipe.finish()                                 # the caller's stream waits for the last gather

    With the persistent SIREN engines, leave a few SMs to the NCCL kernels: `reserve_sms(4)` (undo with `reserve_sms(0)`).
    """

    def __init__(self, n_local, width, chunks=8, device=None, group=None, dtype=torch.float32):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_local, self.width = int(n_local), int(width)
        self.chunks = max(1, min(int(chunks), max(1, self.n_local)))
        self.m = (self.n_local + self.chunks - 1) // self.chunks
        self.ranges = [(min(self.n_local, c * self.m), min(self.n_local, (c + 1) * self.m)) for c in range(self.chunks)]
        dev = torch.device(device) if device is not None else torch.device("cpu")
        self.device = dev
        self.staging = torch.zeros((self.chunks, self.m, self.width), dtype=dtype, device=dev)
        self.gathered = (torch.empty((self.chunks, self.world, self.m, self.width), dtype=dtype, device=dev)
                         if self.world > 1 else self.staging.view(self.chunks, 1, self.m, self.width))
        self.cuda = dev.type == "cuda"
        if self.cuda and self.world > 1:
            self.comm = torch.cuda.Stream(dev)
            self.events = [torch.cuda.Event() for _ in range(self.chunks)]

    def slot(self, c):
        """Staging rows of piece c ([hi - lo, width] view): a producer may write them in place and then call gather(c)."""
        lo, hi = self.ranges[c]
        return self.staging[c, :hi - lo]

    def gather(self, c):
        """Start the all-gather of piece c (its staging rows were written on the current stream)."""
        if self.world == 1:
            return
        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            self.events[c].record(cur)
            self.comm.wait_event(self.events[c])
            with torch.cuda.stream(self.comm):
                dist.all_gather_into_tensor(self.gathered[c].view(self.world * self.m, self.width), self.staging[c], group=self.group)
        else:
            dist.all_gather_into_tensor(self.gathered[c].view(self.world * self.m, self.width), self.staging[c], group=self.group)

    def submit(self, c, *columns):
        """Pack the column blocks (each [rows, k_i], sum k_i = width) of piece c and start its all-gather."""
        lo, hi = self.ranges[c]
        rows = hi - lo
        o = 0
        for col in columns:
            col = col.reshape(rows, -1)
            self.staging[c, :rows, o:o + col.shape[1]] = col
            o += col.shape[1]
        if o != self.width:
            raise RuntimeError(f"PipelinedAllGather.submit: got {o} columns, expected {self.width}")
        self.gather(c)

    def finish(self):
        if self.cuda and self.world > 1:
            torch.cuda.current_stream(self.device).wait_stream(self.comm)
        return self.gathered

    def rows_of_rank(self, r):
        return self.gathered[:, r].reshape(self.chunks * self.m, self.width)[:self.n_local]


class FusedRowGather:
    """The all-gather of the sharded SIREN query FUSED INTO THE QUERY KERNEL.

    Every rank owns `n_local` points; the result is `self.gathered` [world, n_local, 4] = the (value, grad) rows of every
    rank, identical on all ranks.  The buffer is symmetric memory (torch.distributed._symmetric_memory: every rank's copy
    is mapped into every process).  `query()` launches the reverse-mode engine with the peer mappings: its epilogue stores
    each finished row into ALL ranks' buffers over NVLink -- through ONE multimem.st to the NVLink multicast address when
    the fabric supports it (the NVSwitch replicates), otherwise through one store per peer.  There is no communication
    kernel: no SM is reserved, nothing runs after the last tile except a cross-rank barrier.

        fg = FusedRowGather(n_local, device)
        fg.begin()                       # barrier: every rank is done reading the previous result
        fg.query(net, coords_local)      # one launch (or several with row_offset) over the local shard
        fg.finish()                      # barrier: every rank's rows have landed everywhere
        fg.gathered[r]                   # rows of rank r

    mode: "multicast" | "p2p" (what is in use), chosen by `prefer`.  Every rank must pass the same `n_local` (pad the last
    shard): the buffers are symmetric, row r of rank k lives at the same offset everywhere."""

    def __init__(self, n_local, device, group=None, prefer="multicast"):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("FusedRowGather: at most 8 ranks (one NVSwitch domain)")
        self.n_local = int(n_local)
        self.device = torch.device(device)
        self.gathered = symm.empty((self.world, self.n_local, 4), dtype=torch.float32, device=self.device)
        self.hdl = symm.rendezvous(self.gathered, self.group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        mc = 0
        if prefer == "multicast":
            try:
                if self.hdl.has_multicast_support:
                    mc = int(self.hdl.multicast_ptr)
            except Exception:
                mc = 0
        self.mc = mc
        self.mode = "multicast" if mc else "p2p"

    def begin(self):
        self.hdl.barrier(channel=0)

    def query(self, net, coords, row_offset=0):
        off = (self.rank * self.n_local + int(row_offset)) * 16
        net.query_rows_peers(coords, [p + off for p in self.ptrs], (self.mc + off) if self.mc else 0)

    def local_rows(self):
        return self.gathered[self.rank]

    def finish(self):
        self.hdl.barrier(channel=1)

    def rows_of_rank(self, r):
        return self.gathered[r]


class ShardedHostPipeline:
    """The sharded query END TO END for host-resident coordinates: every rank streams ITS shard of the coordinates from
    pinned host memory (H2D of chunk i+1 under the kernel of chunk i), the kernel's epilogue all-gathers the rows into every
    rank's HBM (FusedRowGather), and every rank copies the rows of ONE shard -- `source_rank`, by default its own -- back to
    its host buffer (D2H of chunk i-1 under the kernel of chunk i).  With source_rank != rank the host result of a rank has
    travelled H2D on another GPU -> tensor cores there -> NVLink -> D2H here: the collective is on the timed path.

    A cross-rank barrier follows each chunk's launch (the rows of chunk i are complete on every rank before anybody copies
    them out) and precedes the first one (every rank has finished reading the previous run's rows)."""

    def __init__(self, net, n_local, chunk=1 << 23, device=None, group=None, prefer="multicast"):
        self.net = net
        self.n_local, self.chunk = int(n_local), int(chunk)
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.fg = FusedRowGather(n_local, self.device, group, prefer)
        dev = self.device
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.d_x = [torch.empty((self.chunk, 3), dtype=torch.float32, device=dev) for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_k = [torch.cuda.Event() for _ in range(2)]

    @torch.no_grad()
    def run(self, coords_host, out_host, source_rank=None):
        src = self.fg.rank if source_rank is None else int(source_rank)
        n = coords_host.shape[0]
        if n != self.n_local:
            raise RuntimeError("ShardedHostPipeline.run: coords_host must hold this rank's n_local points")
        cur = torch.cuda.current_stream(self.device)
        self.fg.begin()
        nchunks = (n + self.chunk - 1) // self.chunk
        for i in range(nchunks):
            b = i & 1
            lo, hi = i * self.chunk, min(n, (i + 1) * self.chunk)
            m = hi - lo
            with torch.cuda.stream(self.s_in):
                if i >= 2:
                    self.s_in.wait_event(self.ev_k[b])
                self.d_x[b][:m].copy_(coords_host[lo:hi], non_blocking=True)
                self.ev_in[b].record(self.s_in)
            cur.wait_event(self.ev_in[b])
            self.fg.query(self.net, self.d_x[b][:m], row_offset=lo)
            self.fg.finish()                       # every rank's rows of chunk i have landed everywhere
            self.ev_k[b].record(cur)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_k[b])
                out_host[lo:hi].copy_(self.fg.gathered[src, lo:hi], non_blocking=True)
        self.s_out.synchronize()
        return out_host


def reserve_sms(n):
    """Keep `n` SMs free of the persistent SIREN engines (0 = use all), so that NCCL kernels launched on another stream can
    run next to them (include/inf3dp.h: inf3dp_siren_set_sm_limit)."""
    from . import _lib
    L = _lib.lib()
    if n <= 0:
        _lib.check(L.inf3dp_siren_set_sm_limit(0))
        return
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    _lib.check(L.inf3dp_siren_set_sm_limit(max(2, sms - int(n))))
