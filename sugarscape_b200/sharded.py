"""One simulation sharded over the GPUs of a box by x-slabs (BASELINE S2).

One process per GPU (torchrun).  Grid and agent arrays are "stitched" (include/ssb.h, ssb_fabric_*):
rank r's slab of the map and rank r's range of agent slots live in rank r's HBM, every rank maps the
whole arrays, and kernels reach across the slab boundary over NVLink.  This module is the host
plumbing: passing the arena file descriptors between the ranks (unix sockets, SCM_RIGHTS), laying
the stitched arrays out, re-slotting a host state so that every agent's record sits on the rank that
owns its cell, and combining the per-rank stats reductions.  Scalable RNG mode only.
"""
import ctypes as C
import os
import shutil
import socket
import tempfile
import struct

import numpy as np

from . import layout
from .engine import Engine, EngineError, load_library

FDS_PER_MESSAGE = 64
LOCAL_AGENT_FIELDS = ("order", "free_slots", "dead_list")      # per rank; everything else is indexed by global slot


GRANULE = 2 << 20           # allocation granularity of the stitched arrays (checked against the driver's)


def slab_bounds(width, world):
    """Equal-width slabs: x-lines [b[r], b[r+1]) belong to rank r."""
    return [width * r // world for r in range(world + 1)]


def balanced_slab_bounds(state, world):
    """Slabs with equal POPULATION (agents act, cells only regrow): boundaries at the x-lines where
    the cumulative agent count crosses r / world of the total, rounded to a multiple of `align`
    x-lines so that a slab boundary is also a boundary of 2 MiB pages of the i32 grid arrays
    (then a rank's slab lies in its own HBM to the last line).  Falls back to equal widths for
    empty maps; every slab keeps at least one x-line."""
    W, H = state.width, state.height
    n = int(state.counters[layout.C_N_LIST])
    if n == 0 or world == 1:
        return slab_bounds(W, world)
    align = GRANULE // (H * 4) if GRANULE % (H * 4) == 0 and GRANULE // (H * 4) * 4 * world <= W else 1
    if SWEEP_TILE * 2 * world <= W:        # whole 64-column tiles per slab: the ranks can run the sweep scheduler
        align = align * SWEEP_TILE // np.gcd(align, SWEEP_TILE)
        if align * 2 * world > W:
            align = SWEEP_TILE
    x = state.a_pos[state.a_order[:n]] // H
    cum = np.cumsum(np.bincount(x, minlength=W))
    bounds = [0]
    for r in range(1, world):
        b = int(np.searchsorted(cum, n * r / world, side="left")) + 1
        b = int(round(b / align)) * align
        bounds.append(min(max(b, bounds[-1] + align), W - (world - r) * align))
    return bounds + [W]


SWEEP_TILE = 64          # the sweep scheduler works on 64 x 64 tiles (csrc/ssb_flow_core.cuh): slabs of whole tiles can run it


def sweep_slabs_possible(params, xb):
    """Mirror of the engine's ssb_shard_sweep_possible as far as it can be told before the bind: the rule set of the plain
    lean transaction (csrc/ssb_sweep_core.cuh, lean_supported + lean_mode == 0) and slab boundaries on tile boundaries."""
    heavy = (layout.F_TAGGING | layout.F_TRADE | layout.F_REPRO | layout.F_TAGPREF | layout.F_COMBAT | layout.F_SPICE | layout.F_POLLUTION)
    if params.flags & heavy or params.max_tribes > 25 or params.max_sugar_capacity > 32767:
        return False
    side = min(params.width, params.height)
    if not (0 <= params.max_agent_range <= 7 and 2 * params.max_agent_range + 1 < side):
        return False
    return all(b % SWEEP_TILE == 0 for b in xb[:-1])


def slot_bounds(capacity, world):
    return [capacity * r // world for r in range(world + 1)]


def shard_host_state(state, world, xb=None):
    """Re-slot a freshly built single-GPU HostState for `world` ranks: an agent's record moves to the
    slot range of the rank that owns its cell (list order is kept inside a rank).  xb: slab
    boundaries in x-lines (default: balanced_slab_bounds).  Returns the new global state and, per
    rank, its list of slots; the boundaries are kept as new.slab_bounds."""
    n_list, n_slots = int(state.counters[layout.C_N_LIST]), int(state.counters[layout.C_N_SLOTS])
    if int(state.counters[layout.C_N_FREE]) != 0 or n_list != n_slots:
        raise ValueError("shard_host_state expects a freshly built state (no free slots)")
    if state.ethics is not None or state.ext is not None:
        raise ValueError("decision models, lending, friends, diseases ... exist in the exact RNG mode only: not shardable")
    xb = balanced_slab_bounds(state, world) if xb is None else list(xb)
    sb = slot_bounds(state.capacity, world)
    old = state.a_order[:n_list].astype(np.int64)
    x = state.a_pos[old] // state.height
    owner = np.searchsorted(np.asarray(xb[1:]), x, side="right")
    new = layout.HostState.empty(state.width, state.height, state.capacity, state.kin_capacity)
    for name, _ in layout.GRID_FIELDS:
        setattr(new, name, getattr(state, name).copy())
    new.counters[:] = state.counters
    lists = []
    new_slot = np.empty(n_list, dtype=np.int64)
    for r in range(world):
        idx = np.flatnonzero(owner == r)
        if len(idx) > sb[r + 1] - sb[r]:
            raise MemoryError(f"rank {r}: {len(idx)} agents exceed its {sb[r + 1] - sb[r]} slots")
        new_slot[idx] = sb[r] + np.arange(len(idx))
        lists.append((sb[r] + np.arange(len(idx))).astype(np.int32))
    for name, _, _ in layout.AGENT_FIELDS:
        if name in LOCAL_AGENT_FIELDS:
            continue
        getattr(new, "a_" + name)[new_slot] = getattr(state, "a_" + name)[old]
    new.occ[state.a_pos[old]] = new_slot.astype(np.int32)
    new.slab_bounds = xb
    return new, lists


def chunk_boundaries(ends, itemsize, world):
    """Byte boundaries b[0..world] of the chunks of a stitched array whose rank r should hold the
    elements [ends[r-1], ends[r]): multiples of GRANULE, strictly increasing (a rank whose share
    is smaller than a page still backs one page; small arrays simply live on the first ranks)."""
    b = [0]
    for r in range(world):
        want = int(round(ends[r] * itemsize / GRANULE)) * GRANULE
        if r == world - 1:
            want = -(-ends[r] * itemsize // GRANULE) * GRANULE
        b.append(max(want, b[-1] + GRANULE))
    return b


def rendezvous_directory(rank, dist):
    """A directory only this user can enter, created by rank 0 (tempfile.mkdtemp: mode 0700, unpredictable name) and
    announced to the other ranks through the process group: the unix sockets of exchange_fds live there, not under a
    guessable name in /tmp that anybody could bind first."""
    box = [tempfile.mkdtemp(prefix="ssb_fabric_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def exchange_fds(rank, world, fds, tag, barrier, directory=None):
    """All-to-all of a list of file descriptors per rank over unix sockets (SCM_RIGHTS) in `directory`
    (rendezvous_directory; default: the system's temporary directory, for callers without a process group).
    Returns {rank: [fds]} (this rank's own list included as given)."""
    fds = list(fds)
    if world == 1:
        return {rank: fds}
    directory = directory or tempfile.gettempdir()
    path = lambda r: os.path.join(directory, f"ssb_fabric_{tag}_{r}.sock")
    if os.path.exists(path(rank)):
        os.unlink(path(rank))
    server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    server.bind(path(rank))
    server.listen(world)
    try:
        barrier()                                  # every rank is listening
        for r in range(world):
            if r == rank:
                continue
            with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                c.connect(path(r))
                for i in range(0, len(fds), FDS_PER_MESSAGE):
                    # (a message's ancillary data is a barrier for recvmsg: descriptors of two messages never merge)
                    socket.send_fds(c, [struct.pack("ii", rank, i)], fds[i:i + FDS_PER_MESSAGE])
        out = {rank: fds}
        for _ in range(world - 1):
            conn, _ = server.accept()
            with conn:
                got = []
                while len(got) < len(fds):
                    msg, part, _, _ = socket.recv_fds(conn, 8, FDS_PER_MESSAGE)
                    src, at = struct.unpack("ii", msg)
                    assert at == len(got) and part
                    got += part
            out[src] = got
        fds = out
        barrier()
    finally:
        server.close()
        os.unlink(path(rank))
    return fds


def combine_stats(parts):
    """Per-rank ssb_stats_t -> the box-wide one (sums; extrema; the Lorenz sum is already global)."""
    out = layout.Stats()
    live = [p for p in parts if p.population > 0]
    for field, _ in layout.Stats._fields_:
        vals = [getattr(p, field) for p in parts]
        if field == "timestep":
            v = vals[0]
        elif field in ("rounds",):
            v = max(vals)
        elif field in ("lorenz_weighted_sum", "gini_area"):
            v = vals[0]
        elif field == "max_wealth":
            v = max((p.max_wealth for p in live), default=0.0)
        elif field == "min_wealth":
            v = min((p.min_wealth for p in live), default=0.0)
        elif field == "tribe_first":
            v = type(vals[0])(*[min((x[i] for x in vals if x[i] >= 0), default=-1) for i in range(len(vals[0]))])
        elif hasattr(vals[0], "__len__"):
            v = type(vals[0])(*[sum(x[i] for x in vals) for i in range(len(vals[0]))])
        else:
            v = sum(vals)
        setattr(out, field, v)
    return out


class StitchedArray:
    """A 1-D array in stitched device memory (chunk p backed by rank p's HBM), with the few tensor
    methods the Engine uses.  Copies are synchronous and split at chunk boundaries."""

    def __init__(self, owner, ptr, n, dtype, bounds, offset=0):
        self.owner, self.ptr, self.n, self.dtype, self.bounds = owner, int(ptr), int(n), np.dtype(dtype), bounds
        self.offset = offset          # byte offset of [0] in the mapping (bounds are offsets in the mapping too)

    def data_ptr(self):
        return self.ptr

    def numel(self):
        return self.n

    def element_size(self):
        return self.dtype.itemsize

    def __getitem__(self, sl):
        lo, hi, step = sl.indices(self.n)
        assert step == 1
        return StitchedArray(self.owner, self.ptr + lo * self.dtype.itemsize, max(hi - lo, 0), self.dtype, self.bounds,
                             self.offset + lo * self.dtype.itemsize)

    def _pieces(self):
        """(byte offset inside this view, bytes) pieces that do not straddle a chunk boundary."""
        done, total = 0, self.n * self.dtype.itemsize
        while done < total:
            at = self.offset + done
            nxt = next(b for b in self.bounds if b > at)
            step = min(nxt - at, total - done)
            yield done, step
            done += step

    def _copy(self, host_ptr, kind):
        for off, nbytes in self._pieces():
            dev = self.ptr + off
            args = (dev, host_ptr + off) if kind == 0 else ((host_ptr + off, dev) if kind == 1 else (dev, None))
            self.owner._fcheck(self.owner.L.ssb_fabric_copy(self.owner.fabric, args[0], args[1], nbytes, kind))

    def copy_(self, src):
        arr = np.ascontiguousarray(src.numpy() if hasattr(src, "numpy") else src)
        assert arr.nbytes == self.n * self.dtype.itemsize
        self._copy(arr.ctypes.data, 0)
        return self

    def zero_(self):
        self._copy(0, 2)
        return self

    def cpu(self):
        return self

    def numpy(self):
        out = np.empty(self.n, dtype=self.dtype)
        if self.n:
            self._copy(out.ctypes.data, 1)
        return out


class ShardedEngine(Engine):
    """Rank `rank` of `world` of one simulation.  `host_state` is the GLOBAL state as returned by
    shard_host_state() (identical on every rank), `my_list` this rank's list of slots.  Needs an
    initialised torch.distributed process group (any backend) for the set-up handshakes."""

    def __init__(self, params, host_state, my_list, rank, world, device=0, endow_bits=None, tag_bits=None,
                 mark_shift=None, tag="0"):
        import torch.distributed as dist
        self.rank, self.world, self.my_list, self.dist = rank, world, np.asarray(my_list, dtype=np.int32), dist
        self.fabric = C.c_void_p()
        if mark_shift is None:       # sharded rounds are fewer and larger (128 epochs): the finer 2 x 2 marks win
            mark_shift = min(layout.default_mark_shift(params), 1)
        self.mark_shift = mark_shift
        self.tag = f"{os.environ.get('MASTER_PORT', '0')}_{tag}"
        super().__init__(params, host_state, device, endow_bits, tag_bits)

    # -- set-up --------------------------------------------------------------------------------------
    def _barrier(self):
        self.torch.cuda.synchronize(self.device)
        self.dist.barrier()

    def _fcheck(self, rc):
        if rc != 0:
            raise EngineError("fabric: " + (self.L.ssb_fabric_last_error(self.fabric) or b"?").decode())

    def _upload_all(self, hs):
        t, L, world, rank = self.torch, self.L, self.world, self.rank
        t.cuda.set_device(self.device)
        n_cells, cap = self.width * self.height, self.capacity
        xb = list(getattr(hs, "slab_bounds", None) or slab_bounds(self.width, world))
        sb = slot_bounds(cap, world)
        self.slab_bounds = xb
        self.cell_lo, self.cell_hi = xb[rank] * self.height, xb[rank + 1] * self.height
        self.slot_lo, self.slot_hi = sb[rank], sb[rank + 1]
        local_slots = max(sb[r + 1] - sb[r] for r in range(world))
        sh = self.mark_shift
        wc, hc = self.width >> sh, self.height >> sh
        mark_end = [min(-(-xb[r + 1] // (1 << sh)), wc) * hc for r in range(world)]      # coarse columns, split like the slabs
        mark_end[-1] = wc * hc
        self.mark_lo, self.mark_hi = (mark_end[rank - 1] if rank else 0), mark_end[rank]
        uniform = lambda n: [-(-n // world) * (r + 1) for r in range(world)]
        # (key, elements, dtype, elements each rank should hold) of every stitched array, in mapping order
        plan = [(name, n_cells, dt, [xb[r + 1] * self.height for r in range(world)]) for name, dt in layout.GRID_FIELDS]
        plan += [("a_" + name, cap, dt, sb[1:]) for name, dt, _ in layout.AGENT_FIELDS if name not in LOCAL_AGENT_FIELDS]
        plan += [("_ctrl", (GRANULE // 8) * world, np.uint64, uniform((GRANULE // 8) * world)), ("_marks", wc * hc, np.uint32, mark_end)]
        gather_words = -(-5 * local_slots * 8 // GRANULE) * GRANULE // 8
        plan += [("_gather", gather_words * world, np.uint64, uniform(gather_words * world))]
        # the sweep scheduler's arrays (ssb_set_shard_sweep): descriptors, done bitmap, packed cells -- stitched like the grid,
        # if the slabs consist of whole 64-column tiles (the engine decides after the bind whether it can use them)
        self.sweep_arrays = sweep_slabs_possible(self.params, xb)
        if self.sweep_arrays:
            wpc = (self.height + 63) // 64
            ty4 = (self.height + 3) // 4
            plan += [("_sw_desc", n_cells, np.uint64, [xb[r + 1] * self.height for r in range(world)]),
                     ("_sw_done", self.width * wpc, np.uint64, [xb[r + 1] * wpc for r in range(world)]),
                     ("_sw_pcell", ((self.width + 3) // 4) * ty4 * 16, np.uint64, [min((xb[r + 1] + 3) // 4, (self.width + 3) // 4) * ty4 * 16 for r in range(world)])]
        bounds = {key: chunk_boundaries(ends, np.dtype(dt).itemsize, world) for key, n, dt, ends in plan}
        self._fcheck(L.ssb_fabric_create(rank, world, self.device.index, C.byref(self.fabric)))
        if GRANULE % int(L.ssb_fabric_granularity(self.fabric)):
            raise EngineError("unexpected allocation granularity %d" % L.ssb_fabric_granularity(self.fabric))
        mine = []
        for i, (key, n, dt, ends) in enumerate(plan):          # this rank's chunk of every array
            fd = C.c_int32(-1)
            if L.ssb_fabric_alloc(self.fabric, bounds[key][rank + 1] - bounds[key][rank], C.byref(fd)) != i:
                self._fcheck(-1)
            mine.append(fd.value)
        directory = rendezvous_directory(rank, self.dist)
        try:
            fds = exchange_fds(rank, world, mine, self.tag, self.dist.barrier, directory)
        finally:
            self.dist.barrier()
            if rank == 0:
                shutil.rmtree(directory, ignore_errors=True)
        self.stitched, self.bounds = {}, bounds
        for i, (key, n, dt, ends) in enumerate(plan):
            base = C.c_void_p()
            b = bounds[key]
            per_rank = (C.c_int32 * world)(*[fds[r][i] for r in range(world)])
            sizes = (C.c_uint64 * world)(*[b[r + 1] - b[r] for r in range(world)])
            self._fcheck(L.ssb_fabric_map(self.fabric, i, per_rank, sizes, C.byref(base)))
            view_dt = np.int32 if np.dtype(dt) == np.uint32 else dt       # like Engine._to_device
            self.stitched[key] = StitchedArray(self, base.value, b[-1] // np.dtype(dt).itemsize, view_dt, b)
        for fd in mine:
            os.close(fd)
        # this rank zeroes / fills the memory it backs, then uploads its slab and its slots
        for key, n, dt, ends in plan:
            item, b = np.dtype(dt).itemsize, bounds[key]
            self.stitched[key][b[rank] // item:b[rank + 1] // item].zero_()
        self._barrier()
        for name, _ in layout.GRID_FIELDS:
            self.tensors[name] = self.stitched[name][:n_cells]
            self._put(self.tensors[name], getattr(hs, name), self.cell_lo, self.cell_hi)
        counters = hs.counters.copy()
        counters[layout.C_N_LIST] = len(self.my_list)
        counters[layout.C_N_SLOTS] = self.slot_lo + len(self.my_list)
        counters[layout.C_N_FREE] = 0
        self.tensors["counters"] = self._to_device(counters)
        for name, dt, fill in layout.AGENT_FIELDS:
            if name in LOCAL_AGENT_FIELDS:
                arr = np.full(cap, fill, dtype=dt)
                if name == "order":
                    arr[:len(self.my_list)] = self.my_list
                self.tensors["a_" + name] = self._to_device(arr)
            else:
                self.tensors["a_" + name] = self.stitched["a_" + name][:cap]
                self._put(self.tensors["a_" + name], getattr(hs, "a_" + name), self.slot_lo, self.slot_hi)
        for name in layout.KIN_FIELDS:
            self.tensors[name] = self._to_device(getattr(hs, name))
        self._barrier()

    def _put(self, tensor, host, lo, hi):
        arr = host[lo:hi]
        if arr.dtype == np.uint32:
            arr = arr.view(np.int32)
        tensor[lo:hi].copy_(self.torch.from_numpy(np.ascontiguousarray(arr)))

    def _after_bind(self):
        self.set_mark_shift(self.mark_shift)
        sd = layout.Shard()
        sd.rank, sd.world = self.rank, self.world
        for r in range(layout.MAX_RANKS):
            sd.cell_end[r] = self.slab_bounds[min(r + 1, self.world)] * self.height
        sd.slot_lo, sd.slot_hi, sd.mark_lo, sd.mark_hi = self.slot_lo, self.slot_hi, self.mark_lo, self.mark_hi
        sd.ctrl, sd.ctrl_stride = self.stitched["_ctrl"].data_ptr(), self.bounds["_ctrl"][1]
        sd.marks = self.stitched["_marks"].data_ptr()
        sd.gather, sd.gather_stride = self.stitched["_gather"].data_ptr(), self.bounds["_gather"][1]
        self._check(self.L.ssb_set_shard(self.handle, C.byref(sd)))
        self._barrier()
        # every rank reaches the same verdict (same parameters, same slabs): all of them switch to the sweep, or none
        if self.sweep_arrays and os.environ.get("SSB_SCHEDULER", "") != "rounds" and self.L.ssb_shard_sweep_possible(self.handle):
            self._check(self.L.ssb_set_shard_sweep(self.handle, self.stitched["_sw_desc"].data_ptr(), self.stitched["_sw_done"].data_ptr(),
                                                   self.stitched["_sw_pcell"].data_ptr()))
            self._barrier()

    def close(self):
        if getattr(self, "handle", None):
            self.torch.cuda.synchronize(self.device)
            try:
                self.dist.barrier()                # nobody unmaps memory a peer's kernel may still touch
            except Exception:
                pass
            self.L.ssb_destroy(self.handle)
            self.handle = None
            self.tensors.clear()
            self.stitched.clear()
            if self.fabric:
                self.L.ssb_fabric_destroy(self.fabric)
                self.fabric = None

    # -- box-wide views ----------------------------------------------------------------------------
    def broken(self):
        return bool(self.L.ssb_shard_broken(self.handle))

    def gather_lists(self):
        """Every rank's list of slots, concatenated in rank order (same result on every rank)."""
        n = int(self.counters()[layout.C_N_LIST])
        mine = self.tensors["a_order"][:n].cpu().numpy()
        parts = [None] * self.world
        self.dist.all_gather_object(parts, mine)
        return np.concatenate(parts)

    def download_global(self, fields=None):
        """The box-wide state as one HostState (the stitched arrays are visible from every rank);
        a_order / N_LIST describe the union of the ranks' lists."""
        self._barrier()
        hs = self.download(fields)
        order = self.gather_lists()
        hs.a_order = np.full(self.capacity, -1, dtype=np.int32)
        hs.a_order[:len(order)] = order
        counters = self.counters().copy()
        counters[layout.C_N_LIST] = len(order)
        hs.counters = counters
        self._barrier()
        return hs

    def global_stats(self, t=None):
        mine = self.stats() if t is None else self.stats_history(t)
        parts = [None] * self.world
        self.dist.all_gather_object(parts, bytes(mine))
        return combine_stats([layout.Stats.from_buffer_copy(b) for b in parts])
