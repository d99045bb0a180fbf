#!/usr/bin/env python3
"""Dry-run GPU-marked tests on a machine without a GPU by swapping the CUDA engine for an adapter
over the CPU oracle.  It checks the TEST code (fixtures, invariants, bookkeeping), not the engine:
a development aid for tests that cannot be run until the next GPU slot.

    python tools/run_gpu_tests_on_oracle.py tests/test_gpu_parity.py::test_state_invariants_at_full_size ...
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pytest  # noqa: E402

import sugarscape_b200.engine as engine_mod  # noqa: E402
from oracle_binding import Oracle  # noqa: E402
from sugarscape_b200 import layout  # noqa: E402


class OracleEngine:
    """The subset of sugarscape_b200.engine.Engine the tests use, on top of oracle/ss_oracle.c."""

    def __init__(self, params, host_state, device=0, endow_bits=None, tag_bits=None):
        self.params, self.state = params, host_state
        self.width, self.height, self.capacity = params.width, params.height, host_state.capacity
        self.orc = Oracle(params, host_state, endow_bits, tag_bits)
        self.t_stats = 0
        self.replacing = False
        self.tensors = {}

    # the oracle IS the sequential order: every scheduler name is accepted and means the same
    def set_scheduler(self, name):
        self._scheduler = name

    def scheduler(self):
        return getattr(self, "_scheduler", "rounds")

    def set_mark_shift(self, shift):
        pass

    def set_epochs(self, epochs):
        pass

    def uses_lean_transaction(self):
        return False

    def set_mt_from_python(self):
        self.orc.set_mt_from_python()

    def push_mt_to_python(self):
        self.orc.push_mt_to_python()

    def mt_state(self):
        return self.orc.mt_state()

    def bind_endowments(self, endowments):
        self.orc.bind_endowments(endowments)
        self.replacing = True

    def env_step(self, t):
        self.orc.env_step(t)

    def agent_step(self, t, prev_mean_wealth=0.0):
        self.orc.agent_step(t, prev_mean_wealth)

    def compact(self, t):
        self.orc.compact(t)

    def replace(self, t):
        self.orc.replace(t)

    def reduce_stats(self, t):
        self.t_stats = t

    def step(self, t, prev_mean_wealth=0.0):
        self.env_step(t); self.agent_step(t, prev_mean_wealth); self.compact(t)
        if self.replacing and self.params.rng_mode == layout.RNG_SCALABLE:
            self.replace(t)
        self.t_stats = t

    def stats(self):
        return self.orc.stats(self.t_stats)

    def download(self, fields=None):
        return self.state

    def upload(self, hs, fields):
        pass

    def counters(self):
        return self.state.counters

    def close(self):
        pass


if __name__ == "__main__":
    engine_mod.Engine = OracleEngine
    sys.exit(pytest.main(["-x", "-q", "-p", "no:cacheprovider"] + sys.argv[1:]))
