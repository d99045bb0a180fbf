#!/usr/bin/env python3
"""Dry-run GPU-marked mode E tests without a GPU on the HOST-COMPILED DEVICE SOURCE (tests/hostemu): the
engine adapter steps agents with the same ssb_rules.cuh text nvcc compiles, everything else with the CPU
oracle.  A development aid (checks test code and rule source before a GPU slot), never a product path.

    python tools/run_gpu_tests_on_hostemu.py tests/test_zz_gpu_decision_models.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import pytest  # noqa: E402

import sugarscape_b200.engine as engine_mod  # noqa: E402
from hostemu_binding import HostEmu  # noqa: E402
from run_gpu_tests_on_oracle import OracleEngine  # noqa: E402


class HostEmuEngine(OracleEngine):
    _ext_stats = None

    def __init__(self, params, host_state, device=0, endow_bits=None, tag_bits=None):
        super().__init__(params, host_state, device, endow_bits, tag_bits)
        self.orc = HostEmu(params, host_state, endow_bits, tag_bits)
        self.ethics = host_state.ethics
        self.ext = host_state.ext

    def ethics_stats(self):
        return self.orc.ethics_stats() if self.ethics is not None else None

    def reduce_stats(self, t):          # like ssb_reduce_stats: the extension reductions run with the regular ones
        super().reduce_stats(t)
        self._ext_stats = self.orc.ext_stats(t)

    def ext_stats(self):
        return self._ext_stats

    def agent_records(self):
        return self.orc.agent_records()

    def group_stats(self):
        return self.orc.group_stats(self.t_stats)

    def sync(self):
        pass


if __name__ == "__main__":
    engine_mod.Engine = HostEmuEngine
    sys.exit(pytest.main(["-x", "-q", "-p", "no:cacheprovider", "--runxfail"] + sys.argv[1:]))
