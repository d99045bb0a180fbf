"""Runs at the examples' SHIPPED length against the unmodified reference (tests/golden/make_golden_long.py: state
digests of every 50th step of 1000).  What the 200-step fixtures cannot show: the kin pool (children / mates lists of
the initiating parent, agent.py:521-527) is a bump allocator of 4 x capacity entries -- without recycling the entries of
removed agents it overflows around step 600 of these runs and children are silently no longer recorded (family
happiness and inheritance then diverge).  The oracle, the host-compiled device source and (on a GPU) the CUDA engine
must stay on the reference's trajectory to the last step, with the pool far from its end."""
import gzip
import json
import os

import pytest

import parity_util as pu
from hostemu_binding import HostEmu
from oracle_binding import Oracle
from sugarscape_b200 import layout

with gzip.open(os.path.join(os.path.dirname(__file__), "golden", "long_runs.json.gz"), "rt") as f:
    LONG = json.load(f)


def _replay(sim, c, state, case, mt_state):
    gold = {rec["t"]: rec for rec in case["steps"]}
    rs_mean_wealth = None
    from sugarscape_b200 import init, stats
    sb = stats.StatsBuilder(c, init.equator(c))
    rs = sb.update(sim.stats(0), 0)
    for t in range(1, case["timesteps"] + 1):
        sim.env_step(t)
        sim.agent_step(t, rs["meanWealth"])
        sim.compact(t)
        rs = sb.update(sim.stats(t), t, 0)
        assert int(state.counters[layout.C_OVERFLOW]) == 0, f"overflow code {state.counters[layout.C_OVERFLOW]} at t={t}"
        if t in gold:
            d, snap = pu.state_digests(state, *mt_state())
            assert len(snap["a_id"]) == gold[t]["population"], f"population differs at t={t}"
            assert d == gold[t]["digests"], f"state differs from the reference at t={t}"
    # the pool recycles: it never needed more than a fraction of its capacity
    assert int(state.counters[layout.C_N_KIN]) < state.kin_capacity // 2


@pytest.mark.parametrize("name", sorted(LONG))
def test_oracle_follows_the_reference_for_the_shipped_run_length(name):
    case = LONG[name]
    c, state, pop, params, endow, tags = pu.build(name, overrides={"timesteps": case["timesteps"]})
    orc = Oracle(params, state, endow, tags)
    orc.set_mt_from_python()
    _replay(orc, c, state, case, orc.mt_state)


@pytest.mark.parametrize("name", sorted(LONG))
def test_device_rule_source_follows_the_reference_for_the_shipped_run_length(name):
    case = LONG[name]
    c, state, pop, params, endow, tags = pu.build(name, overrides={"timesteps": case["timesteps"]})
    emu = HostEmu(params, state, endow, tags)
    emu.set_mt_from_python()
    _replay(emu, c, state, case, emu.mt_state)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(LONG))
def test_engine_follows_the_reference_for_the_shipped_run_length(name):
    """The CUDA engine through the drop-in driver's own stepping calls (Engine.step reads the stats every step, which
    raises on any overflow code)."""
    from sugarscape_b200.engine import Engine
    case = LONG[name]
    c, state, pop, params, endow, tags = pu.build(name, overrides={"timesteps": case["timesteps"]})
    eng = Engine(params, state, 0, endow, tags)
    eng.set_mt_from_python()
    gold = {rec["t"]: rec for rec in case["steps"]}
    from sugarscape_b200 import init, stats
    sb = stats.StatsBuilder(c, init.equator(c))
    eng.reduce_stats(0)
    rs = sb.update(eng.stats(), 0)
    for t in range(1, case["timesteps"] + 1):
        eng.step(t, rs["meanWealth"])
        rs = sb.update(eng.stats(), t, 0)          # (raises EngineError on an overflow code)
        if t in gold:
            d, snap = pu.state_digests(eng.download(), *eng.mt_state())
            assert d == gold[t]["digests"], f"state differs from the reference at t={t}"
    counters = eng.counters()
    assert int(counters[layout.C_N_KIN]) < eng.kin_capacity // 2
    eng.close()
