"""The static conflict DAG of the mode S agent step (sugarscape_b200/csrc/ssb_flow_core.cuh) under the CPU suite.

The dataflow scheduler of the CUDA engine rests on one claim: executing the agents in ANY topological order of
"A -> B iff priority(A) < priority(B) and their footprints intersect" equals the sequential priority order of
the mode S definition (the reference's random.shuffle + agent loop, sugarscape.py:400-407).  tests/hostemu
compiles the engine's own conflict predicate, staged-tile scan and transaction source with g++ and executes the
DAG depth first / in a pseudo-random order -- as far from the priority order as the DAG allows.  The results
must equal (a) the fixtures generated from the unmodified reference with its randomness redirected
(tests/golden/make_golden_scalable.py) and (b) the sequential oracle, bit for bit."""
import numpy as np
import pytest

import parity_util as pu
from hostemu_binding import HostEmuFlow, lib
from oracle_binding import Oracle
from sugarscape_b200 import layout

SCALABLE_GOLDEN = pu.load_scalable_golden()


@pytest.mark.parametrize("order_mode", [1, 2])
@pytest.mark.parametrize("name", sorted(SCALABLE_GOLDEN))
def test_any_topological_order_of_the_conflict_dag_matches_patched_reference(name, order_mode):
    case = SCALABLE_GOLDEN[name]
    c, state, pop, params, endow, tags = pu.build(name, layout.RNG_SCALABLE, case["overrides"])
    emu = HostEmuFlow(params, state, endow, tags, order_mode)
    emu.bind_endowments(pop.endowments)
    edges = 0
    for rec in case["steps"]:
        t = rec["t"]
        emu.env_step(t)
        emu.agent_step(t, 0.0)
        edges += int(emu.dag[1])
        emu.compact(t)
        emu.replace(t)
        d, snap = pu.state_digests_by_id(state)
        assert len(snap["a_id"]) == rec["population"], f"population differs at t={t}"
        for k in ("agents_int", "agents_f64", "agents_tags", "grid_int", "pollution"):
            assert d[k] == rec["digests"][k], f"{k} differs from the patched reference at t={t} (order mode {order_mode})"
    assert edges > 0


BIG_CASES = [
    # S2 rules (combat_limited + replacement) and S1 rules (reproduction) on maps larger than one tile, with the
    # torus seam inside tiles (150 is not a multiple of 64) and dense quadrants
    ("combat_limited", {"environmentWidth": 150, "environmentHeight": 200, "startingAgents": 3000, "agentReplacements": 3000,
                        "agentMaxAge": [5, 30], "agentMovement": [6, 6]}, 25),
    ("combat_unlimited", {"environmentWidth": 130, "environmentHeight": 130, "startingAgents": 2500, "agentMovement": [6, 6],
                          "agentAggressionFactor": [0, 2]}, 25),
    ("reproduction_basic", {"environmentWidth": 140, "environmentHeight": 100, "startingAgents": 1500, "agentMovement": [6, 6],
                            "environmentSugarPeaks": [[70, 30, 4], [30, 70, 4], [20, 20, 4], [80, 80, 4]]}, 30),
    ("trade_replacement", {"environmentWidth": 100, "environmentHeight": 70, "startingAgents": 900, "agentReplacements": 900,
                           "agentMovement": [6, 6]}, 25),
    ("cultural_tagging", {"environmentWidth": 90, "environmentHeight": 90, "startingAgents": 1200, "agentMovement": [4, 4]}, 25),
    ("pollution_formation", {"environmentWidth": 110, "environmentHeight": 110, "startingAgents": 1000, "agentMovement": [6, 6]}, 25),
]


@pytest.mark.parametrize("name,overrides,steps", BIG_CASES)
def test_conflict_dag_equals_sequential_oracle_across_tiles(name, overrides, steps):
    c, state, pop, params, endow, tags = pu.build(name, layout.RNG_SCALABLE, overrides)
    ostate = state.copy()
    emu = HostEmuFlow(params, state, endow, tags, order_mode=1)
    orc = Oracle(params, ostate, endow, tags)
    if c["agentReplacements"] > 0:
        emu.bind_endowments(pop.endowments)
        orc.bind_endowments(pop.endowments)
    deepest = 0
    for t in range(1, steps + 1):
        for sim in (emu, orc):
            sim.env_step(t); sim.agent_step(t, 0.0); sim.compact(t)
            if c["agentReplacements"] > 0:
                sim.replace(t)
        deepest = max(deepest, int(emu.dag[2]))
        d_emu, s_emu = pu.state_digests_by_id(state)
        d_orc, s_orc = pu.state_digests_by_id(ostate)
        for k in ("agents_int", "agents_f64", "agents_tags", "grid_int", "pollution"):
            assert d_emu[k] == d_orc[k], f"{name}: {k} differs from the sequential oracle at t={t}"
    assert deepest > 3, "the case never produced a dependency chain"


def test_conflict_predicate_is_exact_and_symmetric():
    """flow_conflict against a brute-force footprint intersection on a small board; flow_column_reach bounds it."""
    rng = np.random.default_rng(5)

    def cross(x, y, r):
        return {(x, y)} | {(x + d, y) for d in range(-r, r + 1)} | {(x, y + d) for d in range(-r, r + 1)}

    def dilate(cells, k):
        out = set(cells)
        for _ in range(k):
            out |= {(x + dx, y + dy) for (x, y) in out for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1))}
        return out

    import ctypes as C
    L = lib()
    L.emu_flow_conflict.argtypes = [C.c_int32] * 6
    L.emu_flow_column_reach.argtypes = [C.c_int32] * 5
    for _ in range(600):
        ra, rb, ka, kb = int(rng.integers(0, 7)), int(rng.integers(0, 7)), int(rng.integers(0, 3)), int(rng.integers(0, 3))
        dx, dy = int(rng.integers(-18, 19)), int(rng.integers(-18, 19))
        truth = bool(dilate(cross(0, 0, ra), ka) & dilate(cross(dx, dy, rb), kb))
        assert bool(L.emu_flow_conflict(dx, dy, ra, rb, ka, kb)) == truth, (dx, dy, ra, rb, ka, kb)
        assert bool(L.emu_flow_conflict(-dx, -dy, rb, ra, kb, ka)) == truth
        if truth:
            assert abs(dy) <= L.emu_flow_column_reach(abs(dx), ra, ka, max(rb, 6), max(kb, 2))
