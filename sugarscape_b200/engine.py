"""ctypes binding of the CUDA engine (libssb.so, include/ssb.h).

PyTorch owns every simulation buffer (device tensors); the library receives raw device pointers.
There is NO CPU fallback: if the CUDA extension or a CUDA device is missing this module raises.
"""
import ctypes as C
import os
import random

import numpy as np

from . import build as _build
from . import layout

_LIB = None


class EngineError(RuntimeError):
    pass


class _Uploader:
    """Host -> device copies of large pageable arrays at PCIe speed: worker threads copy chunks into a small ring of pinned
    staging buffers (numpy releases the GIL), each chunk is DMAed asynchronously on a side stream while the next ones
    are being staged.  (A plain tensor.to(device) of pageable memory runs at a fraction of the link rate, and the
    whole initial state of a large map -- 20 GB for the S2 shape -- goes through here.)"""
    THRESHOLD = 32 << 20
    CHUNK = 32 << 20
    RING = 8
    WORKERS = 4
    _instances = {}

    @classmethod
    def get(cls, torch, device):
        key = (device.type, device.index)
        if key not in cls._instances:
            cls._instances[key] = cls(torch, device)
        return cls._instances[key]

    def __init__(self, torch, device):
        from concurrent.futures import ThreadPoolExecutor
        self.torch, self.device = torch, device
        self.pinned = [torch.empty(self.CHUNK, dtype=torch.uint8, pin_memory=True) for _ in range(self.RING)]
        self.views = [b.numpy() for b in self.pinned]
        self.events = [None] * self.RING
        self.stream = torch.cuda.Stream(device)
        self.pool = ThreadPoolExecutor(self.WORKERS)

    def upload(self, arr):
        t = self.torch
        src = arr.reshape(-1).view(np.uint8)
        out = t.empty(arr.shape, dtype=t.from_numpy(arr[:0]).dtype, device=self.device)
        dst = out.view(t.uint8).reshape(-1)
        staged = []

        def issue():
            fut, k, off, n = staged.pop(0)
            fut.result()
            with t.cuda.stream(self.stream):
                dst[off:off + n].copy_(self.pinned[k][:n], non_blocking=True)
                ev = t.cuda.Event()
                ev.record(self.stream)
            self.events[k] = ev

        for idx, off in enumerate(range(0, src.nbytes, self.CHUNK)):
            k = idx % self.RING
            while any(item[1] == k for item in staged):      # the ring wrapped around: the buffer is still queued
                issue()
            if self.events[k] is not None:
                self.events[k].synchronize()
                self.events[k] = None
            n = min(self.CHUNK, src.nbytes - off)
            staged.append((self.pool.submit(np.copyto, self.views[k][:n], src[off:off + n]), k, off, n))
            while len(staged) > self.WORKERS:
                issue()
        while staged:
            issue()
        self.stream.synchronize()
        return out

    def download(self, tensor):
        """Device -> host of a large tensor through the same ring: DMA into pinned chunks, worker threads copy them out."""
        t = self.torch
        t.cuda.current_stream(self.device).synchronize()
        src = tensor.contiguous().view(t.uint8).reshape(-1)
        out = np.empty(tensor.shape, dtype=t.empty(0, dtype=tensor.dtype).numpy().dtype)
        dst = out.reshape(-1).view(np.uint8)
        copies = [None] * self.RING
        for idx, off in enumerate(range(0, dst.nbytes, self.CHUNK)):
            k = idx % self.RING
            if copies[k] is not None:
                copies[k].result()
            if self.events[k] is not None:
                self.events[k].synchronize()
                self.events[k] = None
            n = min(self.CHUNK, dst.nbytes - off)
            with t.cuda.stream(self.stream):
                self.pinned[k][:n].copy_(src[off:off + n], non_blocking=True)
                ev = t.cuda.Event()
                ev.record(self.stream)

            def finish(ev=ev, k=k, off=off, n=n):
                ev.synchronize()
                np.copyto(dst[off:off + n], self.views[k][:n])
            copies[k] = self.pool.submit(finish)
        for c in copies:
            if c is not None:
                c.result()
        return out


def load_library():
    """Load libssb.so (built in-tree by sugarscape_b200/build.py); fail loudly if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.OUT
    if not os.path.exists(path):
        raise EngineError(f"CUDA engine {path} is missing: run `python -m sugarscape_b200.build` "
                          "(needs nvcc); this engine has no CPU fallback")
    L = C.CDLL(path)
    P, G, A, S = (C.POINTER(layout.Params), C.POINTER(layout.Grid), C.POINTER(layout.Agents),
                  C.POINTER(layout.Stats))
    E = C.c_void_p
    L.ssb_abi_version.restype = C.c_int
    for name in ("ssb_sizeof_params", "ssb_sizeof_grid", "ssb_sizeof_agents", "ssb_sizeof_stats"):
        getattr(L, name).restype = C.c_int
    L.ssb_create.argtypes = [P, C.c_int, C.POINTER(E)]
    L.ssb_bind.argtypes = [E, G, A]
    L.ssb_destroy.argtypes = [E]
    L.ssb_destroy.restype = None
    L.ssb_last_error.argtypes = [E]
    L.ssb_last_error.restype = C.c_char_p
    L.ssb_set_mt_state.argtypes = [E, C.c_void_p, C.c_int32]
    L.ssb_get_mt_state.argtypes = [E, C.c_void_p, C.POINTER(C.c_int32)]
    L.ssb_set_step_tables.argtypes = [E, C.c_void_p, C.c_void_p, C.c_int32]
    L.ssb_set_epochs.argtypes = [E, C.c_int32]
    L.ssb_set_scheduler.argtypes = [E, C.c_int32]
    L.ssb_get_scheduler.argtypes = [E]
    L.ssb_uses_lean_transaction.argtypes = [E]
    L.ssb_sweep_diag.argtypes = [E, C.c_void_p]
    L.ssb_agent_phases.argtypes = [E, C.c_void_p]
    L.ssb_shard_sweep_possible.argtypes = [E]
    L.ssb_set_shard_sweep.argtypes = [E, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ssb_set_mark_shift.argtypes = [E, C.c_int32]
    L.ssb_set_l2_persist.argtypes = [E, C.c_void_p, C.c_int32]
    L.ssb_env_step.argtypes = [E, C.c_int32, C.c_void_p]
    L.ssb_agent_step.argtypes = [E, C.c_int32, C.c_double, C.c_void_p]
    L.ssb_compact.argtypes = [E, C.c_int32, C.c_void_p]
    L.ssb_reduce_stats.argtypes = [E, C.c_int32, C.c_void_p]
    F = C.c_void_p
    L.ssb_fabric_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(F)]
    L.ssb_fabric_alloc.argtypes = [F, C.c_uint64, C.POINTER(C.c_int32)]
    L.ssb_fabric_map.argtypes = [F, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
    L.ssb_fabric_copy.argtypes = [F, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32]
    L.ssb_fabric_granularity.argtypes = [F]
    L.ssb_fabric_granularity.restype = C.c_uint64
    L.ssb_fabric_destroy.argtypes = [F]
    L.ssb_fabric_last_error.argtypes = [F]
    L.ssb_fabric_last_error.restype = C.c_char_p
    L.ssb_set_shard.argtypes = [E, C.POINTER(layout.Shard)]
    L.ssb_shard_barrier.argtypes = [E, C.c_void_p]
    L.ssb_migrate.argtypes = [E, C.c_int32, C.c_void_p]
    L.ssb_shard_broken.argtypes = [E]
    L.ssb_bind_endowments.argtypes = [E, C.POINTER(layout.Endowments)]
    L.ssb_replace.argtypes = [E, C.c_int32, C.c_void_p]
    L.ssb_step.argtypes = [E, C.c_int32, C.c_double, C.c_void_p]
    L.ssb_stats.argtypes = [E, S]
    L.ssb_stats_history.argtypes = [E, C.c_int32, S]
    L.ssb_sync.argtypes = [E, C.c_void_p]
    L.ssb_bind_ethics.argtypes = [E, C.POINTER(layout.Ethics)]
    L.ssb_ethics_stats.argtypes = [E, C.POINTER(layout.EthicsStats)]
    L.ssb_sizeof_ethics.restype = C.c_int
    L.ssb_bind_ext.argtypes = [E, C.POINTER(layout.Ext)]
    L.ssb_ext_stats.argtypes = [E, C.POINTER(layout.ExtStats)]
    L.ssb_sizeof_ext.restype = C.c_int
    L.ssb_group_stats.argtypes = [E, C.POINTER(layout.GroupStats * 3)]
    L.ssb_sizeof_group_stats.restype = C.c_int
    L.ssb_pollution_buffer.argtypes = [E]
    L.ssb_pollution_buffer.restype = C.c_int
    L.ssb_round_trace.argtypes = [E, C.c_int32, C.c_void_p, C.c_int32]
    L.ssb_test_pow.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    L.ssb_launch_count.argtypes = [E]
    L.ssb_launch_count.restype = C.c_int64
    L.ssb_enable_timing.argtypes = [E, C.c_int32]
    L.ssb_timing.argtypes = [E] + [C.POINTER(C.c_float)] * 4
    if L.ssb_abi_version() != layout.ABI_VERSION:
        raise EngineError("libssb.so ABI version does not match sugarscape_b200.layout")
    for name, struct in (("params", layout.Params), ("grid", layout.Grid), ("agents", layout.Agents),
                         ("stats", layout.Stats), ("ethics", layout.Ethics), ("ext", layout.Ext),
                         ("group_stats", layout.GroupStats)):
        if getattr(L, "ssb_sizeof_" + name)() != C.sizeof(struct):
            raise EngineError(f"struct ssb_{name}_t: library and ctypes mirror disagree on size")
    _LIB = L
    return L


EXPORTED_SYMBOLS = (
    "ssb_abi_version", "ssb_create", "ssb_bind", "ssb_destroy", "ssb_last_error",
    "ssb_set_mt_state", "ssb_get_mt_state", "ssb_set_step_tables", "ssb_set_scheduler", "ssb_get_scheduler", "ssb_uses_lean_transaction", "ssb_sweep_diag", "ssb_agent_phases", "ssb_shard_sweep_possible", "ssb_set_shard_sweep", "ssb_set_epochs", "ssb_set_mark_shift", "ssb_set_l2_persist",
    "ssb_sizeof_params", "ssb_sizeof_grid", "ssb_sizeof_agents", "ssb_sizeof_stats",
    "ssb_env_step", "ssb_agent_step", "ssb_compact", "ssb_bind_endowments", "ssb_replace", "ssb_reduce_stats", "ssb_step", "ssb_stats", "ssb_stats_history",
    "ssb_fabric_create", "ssb_fabric_alloc", "ssb_fabric_map", "ssb_fabric_copy", "ssb_fabric_granularity",
    "ssb_fabric_destroy", "ssb_fabric_last_error", "ssb_set_shard", "ssb_migrate", "ssb_shard_barrier", "ssb_shard_broken",
    "ssb_sync", "ssb_pollution_buffer", "ssb_test_pow", "ssb_round_trace", "ssb_launch_count", "ssb_timing", "ssb_enable_timing",
    "ssb_bind_ethics", "ssb_ethics_stats", "ssb_sizeof_ethics", "ssb_bind_ext", "ssb_ext_stats", "ssb_sizeof_ext",
    "ssb_group_stats", "ssb_sizeof_group_stats",
)


class Engine:
    """One simulation resident on one GPU."""

    def __init__(self, params, host_state, device=0, endow_bits=None, tag_bits=None):
        import torch
        if not torch.cuda.is_available():
            raise EngineError("no CUDA device visible: the Sugarscape engine runs on the GPU only")
        self.torch = torch
        self.L = load_library()
        self.params = params
        self.device = torch.device("cuda", device)
        self.width, self.height, self.capacity = host_state.width, host_state.height, host_state.capacity
        self.kin_capacity = host_state.kin_capacity
        self.tensors = {}
        if (host_state.ethics is not None or host_state.ext is not None) and params.rng_mode != layout.RNG_EXACT:
            raise EngineError("decision models, lending and friends exist in the exact RNG mode only")
        self._upload_all(host_state)
        self.handle = C.c_void_p()
        self._check(self.L.ssb_create(C.byref(params), device, C.byref(self.handle)))
        g, a = self._structs()
        self._check(self.L.ssb_bind(self.handle, C.byref(g), C.byref(a)))
        self._after_bind()
        self.ethics = host_state.ethics          # HostEthics (its arrays live in self.tensors as eth_*) or None
        if self.ethics is not None:
            es = self.ethics.as_struct(lambda arr: None)
            for name, _, _ in layout.ETHICS_FIELDS + layout.ETHICS_HOOD_FIELDS:
                setattr(es, name, self.tensors["eth_" + name].data_ptr())
            self._check(self.L.ssb_bind_ethics(self.handle, C.byref(es)))
        self.ext = host_state.ext                # HostExt (arrays in self.tensors as "ext:<family>.<member>") or None
        if self.ext is not None:
            xs = self.ext.as_struct(lambda arr: None)
            for key in self.ext.arrays:
                fam, member = key.split(".")
                setattr(getattr(xs, fam), member, self.tensors["ext:" + key].data_ptr())
            self._check(self.L.ssb_bind_ext(self.handle, C.byref(xs)))
        # child endowments / tag mismatches are drawn from per-timestep tables (steptables.py): a rule set with births must
        # not step beyond them -- the device would fall back to "every choice from the acting parent" without a word
        self.step_table_horizon = None
        if endow_bits is not None:
            eb = np.ascontiguousarray(endow_bits, dtype=np.uint32)
            tb = np.ascontiguousarray(tag_bits, dtype=np.uint32)
            self._check(self.L.ssb_set_step_tables(self.handle, eb.ctypes.data, tb.ctypes.data, len(eb)))
            self.step_table_horizon = len(eb)
        self._births = bool(params.flags & layout.F_REPRO)

    def _check_horizon(self, t):
        if self._births and (self.step_table_horizon is None or t >= self.step_table_horizon):
            raise EngineError(f"timestep {t} lies beyond the step tables ({self.step_table_horizon} entries): "
                              "build the engine with steptables.step_tables(first, last, ...) that cover the run")

    # -- plumbing ------------------------------------------------------------------------------
    def _after_bind(self):
        pass

    def _check(self, rc):
        if rc != 0:
            msg = self.L.ssb_last_error(self.handle)
            raise EngineError(f"libssb error {rc}: {msg.decode() if msg else '?'}")

    def _to_device(self, arr):
        t = self.torch
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.uint32:
            arr = arr.view(np.int32)
        if arr.nbytes >= _Uploader.THRESHOLD and self.device.type == "cuda":
            return _Uploader.get(t, self.device).upload(arr)
        return t.from_numpy(arr).to(self.device)

    def _upload_all(self, hs):
        for name, _ in layout.GRID_FIELDS:
            self.tensors[name] = self._to_device(getattr(hs, name))
        self.tensors["counters"] = self._to_device(hs.counters)
        for name, _, _ in layout.AGENT_FIELDS:
            self.tensors["a_" + name] = self._to_device(getattr(hs, "a_" + name))
        for name in layout.KIN_FIELDS:
            self.tensors[name] = self._to_device(getattr(hs, name))
        if hs.ethics is not None:
            for name, _, _ in layout.ETHICS_FIELDS + layout.ETHICS_HOOD_FIELDS:
                self.tensors["eth_" + name] = self._to_device(hs.ethics.arrays[name])
        if hs.ext is not None:      # as raw bytes: the families hold uint64 and struct arrays torch has no dtype for
            self._ext_dtype = {key: arr.dtype for key, arr in hs.ext.arrays.items()}
            for key, arr in hs.ext.arrays.items():
                self.tensors["ext:" + key] = self._to_device(np.ascontiguousarray(arr).view(np.uint8))

    def _structs(self):
        g = layout.Grid()
        g.n_cells = self.width * self.height
        for name, _ in layout.GRID_FIELDS:
            setattr(g, name, self.tensors[name].data_ptr())
        a = layout.Agents()
        a.capacity = self.capacity
        a.counters = self.tensors["counters"].data_ptr()
        for name, _, _ in layout.AGENT_FIELDS:
            setattr(a, name, self.tensors["a_" + name].data_ptr())
        a.kin_capacity = self.kin_capacity
        for name in layout.KIN_FIELDS:
            setattr(a, name, self.tensors[name].data_ptr())
        return g, a

    @property
    def stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        if getattr(self, "handle", None):
            self.torch.cuda.synchronize(self.device)
            self.L.ssb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state transfer ------------------------------------------------------------------------
    def download(self, fields=None):
        """Copy the device state back into a HostState (all fields, or the named subset)."""
        hs = layout.HostState(self.width, self.height, self.capacity)
        hs.kin_capacity = self.kin_capacity
        if self.ethics is not None:
            hs.ethics = self.ethics.copy()
        if self.ext is not None:
            hs.ext = self.ext.copy()
        self.torch.cuda.synchronize(self.device)
        swapped = bool(self.L.ssb_pollution_buffer(self.handle))
        for key, tensor in self.tensors.items():
            if fields is not None and key not in fields:
                continue
            if swapped and key in ("pollution", "pollution_flux"):      # the buffers swap roles at diffusion steps
                tensor = self.tensors["pollution_flux" if key == "pollution" else "pollution"]
            if (hasattr(tensor, "is_cuda") and tensor.is_cuda and tensor.numel() * tensor.element_size() >= _Uploader.THRESHOLD):
                arr = _Uploader.get(self.torch, self.device).download(tensor)
            else:
                arr = tensor.cpu().numpy()
            if key == "a_tags":
                arr = arr.view(np.uint32)
            if key.startswith("eth_"):
                hs.ethics.arrays[key[4:]] = arr
            elif key.startswith("ext:"):
                hs.ext.arrays[key[4:]] = arr.view(self._ext_dtype[key[4:]])
            else:
                setattr(hs, key, arr)
        return hs

    def upload(self, hs, fields):
        swapped = bool(self.L.ssb_pollution_buffer(self.handle))
        for key in fields:
            arr = (hs.ethics.arrays[key[4:]] if key.startswith("eth_") else
                   np.ascontiguousarray(hs.ext.arrays[key[4:]]).view(np.uint8) if key.startswith("ext:") else getattr(hs, key))
            if swapped and key in ("pollution", "pollution_flux"):
                key = "pollution_flux" if key == "pollution" else "pollution"
            if arr.dtype == np.uint32:
                arr = arr.view(np.int32)
            self.tensors[key].copy_(self.torch.from_numpy(np.ascontiguousarray(arr)))

    def counters(self):
        return self.tensors["counters"].cpu().numpy()

    # -- RNG (mode E) --------------------------------------------------------------------------
    def set_mt_from_python(self):
        st = random.getstate()[1]
        arr = np.array(st[:624], dtype=np.uint32)
        self._check(self.L.ssb_set_mt_state(self.handle, arr.ctypes.data, st[624]))

    def mt_state(self):
        arr = np.zeros(624, dtype=np.uint32)
        idx = C.c_int32(0)
        self._check(self.L.ssb_get_mt_state(self.handle, arr.ctypes.data, C.byref(idx)))
        return arr, idx.value

    def push_mt_to_python(self):
        arr, idx = self.mt_state()
        random.setstate((3, tuple(int(v) for v in arr) + (idx,), None))

    # -- stepping ------------------------------------------------------------------------------
    def env_step(self, t):
        self._check(self.L.ssb_env_step(self.handle, t, self.stream))

    def agent_step(self, t, prev_mean_wealth=0.0):
        self._check_horizon(t)
        self._check(self.L.ssb_agent_step(self.handle, t, float(prev_mean_wealth), self.stream))

    def compact(self, t):
        self._check(self.L.ssb_compact(self.handle, t, self.stream))

    def bind_endowments(self, endowments):
        """Upload the endowment table replacements draw from (list of dicts from
        init.randomize_endowments, or the struct-of-arrays of layout.endowment_arrays)."""
        arrays = endowments if isinstance(endowments, dict) else layout.endowment_arrays(endowments)
        self.endowment_tensors = {k: self._to_device(v) for k, v in arrays.items()}
        es = layout.endowments_struct(arrays, lambda a: None)
        for name, _ in layout.ENDOWMENT_FIELDS:
            setattr(es, name, self.endowment_tensors[name].data_ptr())
        self._check(self.L.ssb_bind_endowments(self.handle, C.byref(es)))

    def replace(self, t):
        self._check(self.L.ssb_replace(self.handle, t, self.stream))

    def reduce_stats(self, t):
        self._check(self.L.ssb_reduce_stats(self.handle, t, self.stream))

    def step(self, t, prev_mean_wealth=0.0):
        self._check_horizon(t)
        self._check(self.L.ssb_step(self.handle, t, float(prev_mean_wealth), self.stream))

    def stats(self):
        out = layout.Stats()
        self._check(self.L.ssb_stats(self.handle, C.byref(out)))
        return out

    def ethics_stats(self):
        """Decision-model reductions of the last reduce_stats (None without decision models)."""
        if self.ethics is None:
            return None
        out = layout.EthicsStats()
        self._check(self.L.ssb_ethics_stats(self.handle, C.byref(out)))
        return out

    def ext_stats(self):
        """Lending / friends reductions of the last reduce_stats (None without those families); raises when a
        loan pool overflowed."""
        if self.ext is None:
            return None
        if "lending" in self.ext.families:
            if int(self.tensors["ext:lending.counters"].cpu().numpy().view(np.int64)[layout.LC_OVERFLOW]) != 0:
                raise EngineError("loan pool exhausted")
        out = layout.ExtStats()
        self._check(self.L.ssb_ext_stats(self.handle, C.byref(out)))
        return out

    def agent_records(self):
        """The per-agent log records of the last agent step, in action order ([n, layout.AL_FIELDS] float64; None without
        an agentLogfile); resets the device counter for the next step."""
        if self.ext is None or "agentlog" not in self.ext.families:
            return None
        counter = self.tensors["ext:agentlog.n_records"]
        n = int(counter.cpu().numpy().view(np.int64)[0])
        if n >= self.capacity:
            raise EngineError("per-agent log: record buffer exhausted")
        raw = self.tensors["ext:agentlog.records"][:n * layout.AL_FIELDS * 8].cpu().numpy().view(np.float64)
        counter.zero_()
        return raw.reshape(n, layout.AL_FIELDS).copy()

    def group_stats(self):
        """The experimental-group split of the last reduce_stats: layout.GroupStats[3] (None without a group)."""
        if self.ext is None or "group" not in self.ext.families:
            return None
        out = (layout.GroupStats * 3)()
        self._check(self.L.ssb_group_stats(self.handle, C.byref(out)))
        return out

    def stats_history(self, t):
        out = layout.Stats()
        self._check(self.L.ssb_stats_history(self.handle, t, C.byref(out)))
        return out

    def sync(self):
        self._check(self.L.ssb_sync(self.handle, self.stream))

    def round_trace(self, rounds=1024, fetch=True):
        """Enable (first call) and fetch the per-round trace of the last scalable agent step."""
        out = np.zeros((rounds, 6), dtype=np.uint64)
        self._check(self.L.ssb_round_trace(self.handle, rounds, out.ctypes.data if fetch else None, rounds))
        return out

    def set_mark_shift(self, shift):
        self._check(self.L.ssb_set_mark_shift(self.handle, shift))

    def set_l2_persist(self, on=True):
        self._check(self.L.ssb_set_l2_persist(self.handle, self.stream, 1 if on else 0))

    def set_scheduler(self, name):
        """Mode S scheduler: "sweep" (ordered sweep over the static conflict DAG, default on one GPU; lean transaction where
        the rule set allows), "sweep-generic" (the sweep with the generic transaction), "dataflow" (the DAG through ready
        queues) or "rounds" (reservation rounds)."""
        self._check(self.L.ssb_set_scheduler(self.handle, {"rounds": 0, "dataflow": 1, "sweep": 3, "sweep-generic": 4}[name]))

    def scheduler(self):
        return {0: "rounds", 1: "dataflow", 3: "sweep", 4: "sweep-generic"}.get(int(self.L.ssb_get_scheduler(self.handle)), "n/a")

    def sweep_diag(self):
        """{warp iterations, entries held at a readiness check, entries run} of the sweep kernel (SSB_SWEEP_DIAG=1), else None."""
        out = np.zeros(4, dtype=np.uint64)
        if self.L.ssb_sweep_diag(self.handle, out.ctypes.data) != 0:
            return None
        return [int(v) for v in out]

    def uses_lean_transaction(self):
        return bool(self.L.ssb_uses_lean_transaction(self.handle))

    def set_epochs(self, epochs):
        self._check(self.L.ssb_set_epochs(self.handle, epochs))

    def launch_count(self):
        return int(self.L.ssb_launch_count(self.handle))

    def enable_timing(self, on=True):
        self._check(self.L.ssb_enable_timing(self.handle, 1 if on else 0))

    def timing(self):
        vals = [C.c_float(0) for _ in range(4)]
        self._check(self.L.ssb_timing(self.handle, *[C.byref(v) for v in vals]))
        out = {"env_ms": vals[0].value, "agent_ms": vals[1].value, "compact_ms": vals[2].value,
               "stats_ms": vals[3].value}
        sub = np.zeros(5, dtype=np.float32)
        if self.L.ssb_agent_phases(self.handle, sub.ctypes.data) == 0:      # sweep scheduler: the kernels of the agent step
            out.update({"agent_prepare_ms": float(sub[0]), "agent_build_ms": float(sub[1]), "agent_pack_ms": float(sub[2]),
                        "agent_sweep_ms": float(sub[3]), "agent_unpack_ms": float(sub[4])})
        return out
