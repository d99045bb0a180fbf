"""ctypes binding of tests/hostemu/libssb_hostemu.so: the engine's mode E device rule source compiled
for the host as one lane (tests/hostemu/cuda_shim.h).  TEST INFRASTRUCTURE ONLY -- the product has no CPU
path.  `HostEmu` steps a HostState like `oracle_binding.Oracle`, but its agent step is the DEVICE source;
environment step, compaction and reductions stay the oracle's."""
import ctypes as C
import os
import random
import subprocess

import numpy as np

from oracle_binding import Oracle
from sugarscape_b200 import layout

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", HERE])
        L = C.CDLL(os.path.join(HERE, "libssb_hostemu.so"))
        P, G, A = C.POINTER(layout.Params), C.POINTER(layout.Grid), C.POINTER(layout.Agents)
        L.emu_agent_step_exact.argtypes = [P, G, A, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_int32,
                                           C.POINTER(layout.Ethics), C.POINTER(layout.Ext)]
        L.emu_group_stats.argtypes = [A, C.POINTER(layout.Ethics), C.POINTER(layout.Ext), C.c_int32, C.POINTER(layout.GroupStats * 3)]
        L.emu_ext_stats.argtypes = [A, C.POINTER(layout.Ext), C.c_int32, C.POINTER(layout.ExtStats)]
        L.emu_ethics_stats.argtypes = [A, C.POINTER(layout.Ethics), C.POINTER(layout.EthicsStats)]
        assert L.emu_abi_version() == layout.ABI_VERSION
        _LIB = L
    return _LIB


class HostEmu(Oracle):
    def __init__(self, params, state, endow_bits=None, tag_bits=None):
        super().__init__(params, state, endow_bits, tag_bits)
        assert params.rng_mode == layout.RNG_EXACT, "the emulation covers the ordered mode E loop"
        self.E = lib()
        self.mt = np.zeros(625, dtype=np.uint32)
        self.mt[624] = 624

    # the MT19937 stream lives in the engine's layout: 624 words + index
    def set_mt_from_python(self):
        self.mt[:] = np.array(random.getstate()[1], dtype=np.uint32)

    def mt_state(self):
        return self.mt[:624].copy(), int(self.mt[624])

    def agent_step(self, t, prev_mean_wealth=0.0):
        g, a = self.state.as_structs()
        eb = self.endow_bits.ctypes.data if self.endow_bits is not None else None
        tb = self.tag_bits.ctypes.data if self.tag_bits is not None else None
        n = len(self.endow_bits) if self.endow_bits is not None else 0
        eth = C.byref(self.state.ethics.as_struct()) if self.state.ethics is not None else None
        ext = C.byref(self.state.ext.as_struct()) if self.state.ext is not None else None
        rc = self.E.emu_agent_step_exact(C.byref(self.params), C.byref(g), C.byref(a), self.mt.ctypes.data, t,
                                         prev_mean_wealth, eb, tb, n, eth, ext)
        assert rc == 0, f"emu_agent_step_exact failed: {rc}"
        if self.state.ext is not None:
            self.state.ext.check_overflow()

    def agent_records(self):
        """Per-agent log records of the last agent step (like Engine.agent_records)."""
        ext = self.state.ext
        if ext is None or "agentlog" not in ext.families:
            return None
        n = int(ext.arrays["agentlog.n_records"][0])
        raw = ext.arrays["agentlog.records"][:n * layout.AL_FIELDS].reshape(n, layout.AL_FIELDS).copy()
        ext.arrays["agentlog.n_records"][0] = 0
        return raw

    def group_stats(self, t):
        if self.state.ext is None or "group" not in self.state.ext.families:
            return None
        g, a = self.state.as_structs()
        eth = C.byref(self.state.ethics.as_struct()) if self.state.ethics is not None else None
        out = (layout.GroupStats * 3)()
        rc = self.E.emu_group_stats(C.byref(a), eth, C.byref(self.state.ext.as_struct()), t, C.byref(out))
        assert rc == 0
        return out

    def ext_stats(self, t):
        if self.state.ext is None:
            return None
        g, a = self.state.as_structs()
        out = layout.ExtStats()
        rc = self.E.emu_ext_stats(C.byref(a), C.byref(self.state.ext.as_struct()), t, C.byref(out))
        assert rc == 0
        return out

    def ethics_stats(self):
        if self.state.ethics is None:
            return None
        g, a = self.state.as_structs()
        out = layout.EthicsStats()
        rc = self.E.emu_ethics_stats(C.byref(a), C.byref(self.state.ethics.as_struct()), C.byref(out))
        assert rc == 0
        return out


class HostEmuFlow(Oracle):
    """Mode S agent step through the engine's static conflict DAG (csrc/ssb_flow_core.cuh, host-compiled) in an
    adversarial topological order (1 = depth first, 2 = pseudo-random among the ready agents, 0 = FIFO);
    environment step, compaction, replacement and reductions stay the oracle's."""

    def __init__(self, params, state, endow_bits=None, tag_bits=None, order_mode=1):
        super().__init__(params, state, endow_bits, tag_bits)
        assert params.rng_mode == layout.RNG_SCALABLE
        self.E = lib()
        P, G, A = C.POINTER(layout.Params), C.POINTER(layout.Grid), C.POINTER(layout.Agents)
        self.E.emu_agent_step_flow.argtypes = [P, G, A, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_int32,
                                               C.c_void_p, C.c_int32, C.c_void_p]
        self.order_mode = order_mode
        self.born = np.zeros(state.capacity, dtype=np.int32)
        self.dag = np.zeros(4, dtype=np.int64)       # agents, edges, longest path, tiles of the last step

    def agent_step(self, t, prev_mean_wealth=0.0):
        g, a = self.state.as_structs()
        eb = self.endow_bits.ctypes.data if self.endow_bits is not None else None
        tb = self.tag_bits.ctypes.data if self.tag_bits is not None else None
        n = len(self.endow_bits) if self.endow_bits is not None else 0
        rc = self.E.emu_agent_step_flow(C.byref(self.params), C.byref(g), C.byref(a), t, prev_mean_wealth, eb, tb, n,
                                        self.born.ctypes.data, self.order_mode, self.dag.ctypes.data)
        assert rc == 0, f"emu_agent_step_flow failed: {rc}"
        # the engine appends the newborn at compaction; the oracle's compaction expects them on the list
        cn = self.state.counters
        n_born, n_list = int(cn[layout.C_N_BORN]), int(cn[layout.C_N_LIST])
        self.state.a_order[n_list:n_list + n_born] = self.born[:n_born]
        cn[layout.C_N_LIST] = n_list + n_born


class HostEmuSweep(HostEmuFlow):
    """Mode S agent step through the ordered sweep (csrc/ssb_sweep_core.cuh, host-compiled): slabs, gates (near
    predecessors), watermark, done bitmap and -- with lean=True -- the lean transaction on packed records, run in passes
    over the pending entries in an adversarial order (1 = descending priority, 2 = pseudo-random, 0 = ascending list
    index; | 0x100: one priority bucket, so that every predecessor is a near one and the mask lines are exercised)."""

    def __init__(self, params, state, endow_bits=None, tag_bits=None, order_mode=1, lean=False):
        super().__init__(params, state, endow_bits, tag_bits, order_mode)
        P, G, A = C.POINTER(layout.Params), C.POINTER(layout.Grid), C.POINTER(layout.Agents)
        self.E.emu_agent_step_sweep.argtypes = [P, G, A, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_int32,
                                                C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        self.lean = lean
        self.dag = np.zeros(5, dtype=np.int64)       # agents, near predecessors, passes, tiles, mask lines of the last step

    def agent_step(self, t, prev_mean_wealth=0.0):
        g, a = self.state.as_structs()
        eb = self.endow_bits.ctypes.data if self.endow_bits is not None else None
        tb = self.tag_bits.ctypes.data if self.tag_bits is not None else None
        n = len(self.endow_bits) if self.endow_bits is not None else 0
        rc = self.E.emu_agent_step_sweep(C.byref(self.params), C.byref(g), C.byref(a), t, prev_mean_wealth, eb, tb, n,
                                         self.born.ctypes.data, self.order_mode, 1 if self.lean else 0, self.dag.ctypes.data)
        assert rc == 0, f"emu_agent_step_sweep failed: {rc}"
        cn = self.state.counters
        n_born, n_list = int(cn[layout.C_N_BORN]), int(cn[layout.C_N_LIST])
        self.state.a_order[n_list:n_list + n_born] = self.born[:n_born]
        cn[layout.C_N_LIST] = n_list + n_born
