"""The ordered sweep of the mode S agent step (sugarscape_b200/csrc/ssb_sweep_core.cuh) under the CPU suite.

The sweep scheduler of the CUDA engine runs an agent once its PREDECESSORS -- the agents with a smaller priority whose
footprint intersects its own -- have finished: the watermark over the priority buckets vouches for the distant ones,
the agent's gate lists the near ones (found tile by tile in per-slab occupancy bit-planes).  tests/hostemu compiles the
engine's own predecessor scan, gate test and lean transaction (packed records, one thread per agent) with g++ and runs
the agents in passes over the pending entries in descending priority / pseudo-random order -- as far from the
sequential order as the gates allow.  The results must equal (a) the fixtures generated from
the unmodified reference with its randomness redirected (tests/golden/make_golden_scalable.py) and (b) the sequential
oracle, bit for bit."""
import pytest

import parity_util as pu
from hostemu_binding import HostEmuSweep
from oracle_binding import Oracle
from sugarscape_b200 import layout

SCALABLE_GOLDEN = pu.load_scalable_golden()
LEAN_RULES = ("constant_growback", "pollution_formation", "combat_unlimited", "agent_replacement", "combat_limited")
GENERIC_RULES = LEAN_RULES + ("trade_basic", "trade_replacement")     # rule sets with births at range 6: conflict reach 16 > 15 cells


def _replay_golden(name, lean, order_mode):
    case = SCALABLE_GOLDEN[name]
    c, state, pop, params, endow, tags = pu.build(name, layout.RNG_SCALABLE, case["overrides"])
    emu = HostEmuSweep(params, state, endow, tags, order_mode, lean=lean)
    emu.bind_endowments(pop.endowments)
    bits = passes = 0
    for rec in case["steps"]:
        t = rec["t"]
        emu.env_step(t)
        emu.agent_step(t, 0.0)
        bits += int(emu.dag[1]); passes = max(passes, int(emu.dag[2]))
        emu.compact(t)
        emu.replace(t)
        d, snap = pu.state_digests_by_id(state)
        assert len(snap["a_id"]) == rec["population"], f"population differs at t={t}"
        for k in ("agents_int", "agents_f64", "agents_tags", "grid_int", "pollution"):
            assert d[k] == rec["digests"][k], f"{k} differs from the patched reference at t={t} (lean={lean}, order mode {order_mode})"
    assert bits > 0 and passes > 1


@pytest.mark.parametrize("order_mode", [1, 2, 0x101])
@pytest.mark.parametrize("name", LEAN_RULES)
def test_lean_transaction_under_the_sweep_matches_patched_reference(name, order_mode):
    _replay_golden(name, True, order_mode)


@pytest.mark.parametrize("name", GENERIC_RULES)
def test_generic_transaction_under_the_sweep_matches_patched_reference(name):
    _replay_golden(name, False, 1)


BIG_CASES = [
    # S2 rules (combat_limited + replacement) on maps larger than one tile, with the torus seam inside tiles (150 is not a
    # multiple of 64) and dense quadrants; combat with mixed aggression; pollution; and rule sets with dilation 1
    ("combat_limited", {"environmentWidth": 150, "environmentHeight": 200, "startingAgents": 3000, "agentReplacements": 3000,
                        "agentMaxAge": [5, 30], "agentMovement": [6, 6]}, 25, True),
    ("combat_limited", {"environmentWidth": 150, "environmentHeight": 200, "startingAgents": 3000, "agentReplacements": 3000,
                        "agentMaxAge": [5, 30], "agentMovement": [6, 6]}, 10, False),
    ("combat_unlimited", {"environmentWidth": 130, "environmentHeight": 130, "startingAgents": 2500, "agentMovement": [6, 6],
                          "agentAggressionFactor": [0, 2]}, 25, True),
    ("pollution_formation", {"environmentWidth": 110, "environmentHeight": 110, "startingAgents": 1000, "agentMovement": [6, 6]}, 25, True),
    ("seasonal_migration", {"environmentWidth": 70, "environmentHeight": 160, "startingAgents": 1500, "agentMovement": [7, 7], "agentVision": [1, 7]}, 25, True),
    ("trade_replacement", {"environmentWidth": 100, "environmentHeight": 70, "startingAgents": 900, "agentReplacements": 900,
                           "agentMovement": [6, 6]}, 20, False),
    ("cultural_tagging", {"environmentWidth": 90, "environmentHeight": 90, "startingAgents": 1200, "agentMovement": [4, 4], "agentVision": [1, 4]}, 20, False),
]


@pytest.mark.parametrize("order_mode", [1, 0x102])
@pytest.mark.parametrize("name,overrides,steps,lean", BIG_CASES)
def test_sweep_equals_sequential_oracle_across_tiles(name, overrides, steps, lean, order_mode):
    c, state, pop, params, endow, tags = pu.build(name, layout.RNG_SCALABLE, overrides)
    ostate = state.copy()
    emu = HostEmuSweep(params, state, endow, tags, order_mode=order_mode, lean=lean)
    orc = Oracle(params, ostate, endow, tags)
    if c["agentReplacements"] > 0:
        emu.bind_endowments(pop.endowments)
        orc.bind_endowments(pop.endowments)
    deepest = lines = 0
    for t in range(1, steps + 1):
        for sim in (emu, orc):
            sim.env_step(t); sim.agent_step(t, 0.0); sim.compact(t)
            if c["agentReplacements"] > 0:
                sim.replace(t)
        deepest = max(deepest, int(emu.dag[2]))
        lines += int(emu.dag[4])
        d_emu, s_emu = pu.state_digests_by_id(state)
        d_orc, s_orc = pu.state_digests_by_id(ostate)
        for k in ("agents_int", "agents_f64", "agents_tags", "grid_int", "pollution"):
            assert d_emu[k] == d_orc[k], f"{name}: {k} differs from the sequential oracle at t={t}"
    assert deepest > 3, "the case never produced a dependency chain"
    if order_mode & 0x100:
        assert lines > 0, "no agent had more near predecessors than a gate holds: the mask lines were not exercised"
