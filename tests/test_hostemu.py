"""The engine's mode E DEVICE rule source (csrc/ssb_rules.cuh, ssb_exact.cuh), compiled for the host as a
single lane (tests/hostemu), against the fixtures generated from the unmodified reference: every digest
of every step and all 59 log keys of the shipped examples the engine supports.  This puts the transaction
code nvcc compiles for sm_100a under the CPU suite; the `-m gpu` parity tests remain the proof for the
GPU build (lane-parallel ranking, device libm)."""
import pytest

import parity_util as pu
from hostemu_binding import HostEmu
from sugarscape_b200 import init, stats
from test_oracle_golden import SUPPORTED, _close


def run_case(c, state, pop, params, endow, tags, gold, log, setup=None, extra_keys=None):
    emu = HostEmu(params, state, endow, tags)
    if setup:
        setup(emu)
    emu.set_mt_from_python()
    d, _ = pu.state_digests(state, *emu.mt_state())
    assert d == gold[0]["digests"], "t=0 state differs"
    sb = stats.StatsBuilder(c, init.equator(c))
    rs = sb.update(emu.stats(0), 0)
    for t in range(1, len(gold)):
        emu.env_step(t)
        emu.agent_step(t, rs["meanWealth"])
        emu.compact(t)
        replaced = 0
        if c["agentReplacements"] > 0:
            emu.push_mt_to_python()
            replaced = pu.replace_dead_agents(c, state, pop, t)
            emu.set_mt_from_python()
        d, snap = pu.state_digests(state, *emu.mt_state())
        assert d == gold[t]["digests"], f"state differs at t={t}"
        rs = sb.update(emu.stats(t), t, replaced)
        if extra_keys:
            extra_keys(emu, rs, t)
        for k in stats.LOG_KEYS:
            assert _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs[k]} vs {log[t][k]}"
    return emu


@pytest.mark.parametrize("name", SUPPORTED)
def test_device_rule_source_matches_reference_trajectory(name):
    c, state, pop, params, endow, tags = pu.build(name)
    meta, _ = pu.load_golden(name)
    run_case(c, state, pop, params, endow, tags, meta["steps"], meta["log"])
