"""Size-independent invariants of a Sugarscape state (what can be checked at
BASELINE's full sizes, where neither the oracle nor the reference can follow)."""
import numpy as np

from sugarscape_b200 import layout


def check_state(hs, stats=None):
    """Raises AssertionError unless the HostState is self-consistent:
    every listed agent is alive and stands on a cell that points back at it, no two agents share a
    cell, every occupied cell belongs to a listed agent, IDs are unique, holdings are non-negative,
    levels stay within capacity, and (optionally) the stats reduction agrees with the list."""
    n = int(hs.counters[layout.C_N_LIST])
    slots = hs.a_order[:n].astype(np.int64)
    assert len(np.unique(slots)) == n, "a slot is listed twice"
    assert np.all(hs.a_cod[slots] == 0), "a dead agent is still listed"
    pos = hs.a_pos[slots].astype(np.int64)
    assert np.all((pos >= 0) & (pos < hs.width * hs.height)), "an agent stands outside the map"
    assert len(np.unique(pos)) == n, "two agents share a cell"
    assert np.array_equal(hs.occ[pos], slots.astype(hs.occ.dtype)), "occ[] does not point back at the agent on the cell"
    assert int(np.count_nonzero(hs.occ >= 0)) == n, "an occupied cell belongs to no listed agent"
    ids = hs.a_id[slots]
    assert len(np.unique(ids)) == n and np.all(ids >= 0), "agent IDs are not unique"
    assert int(ids.max(initial=-1)) < int(hs.counters[layout.C_NEXT_ID]), "an ID beyond nextAgentID"
    assert np.all(hs.a_sugar[slots] >= 0) and np.all(hs.a_spice[slots] >= 0), "negative holdings on a living agent"
    assert np.all((hs.sugar >= 0) & (hs.sugar <= hs.sugar_cap)), "sugar level outside [0, capacity]"
    assert np.all((hs.spice >= 0) & (hs.spice <= hs.spice_cap)), "spice level outside [0, capacity]"
    assert int(hs.counters[layout.C_OVERFLOW]) == 0, "engine failure flag set"
    if stats is not None:
        assert int(stats.population) == n
        assert int(stats.sum_age) == int(hs.a_age[slots].sum())
        assert int(stats.sum_vision) == int(hs.a_vision[slots].sum())
        wealth = float((hs.a_sugar[slots] + hs.a_spice[slots]).sum())
        assert abs(float(stats.sum_wealth) - wealth) <= 1e-9 * max(1.0, wealth)
    return n


def check_step_accounting(prev_population, stats):
    """population(t) = population(t-1) - deaths + births + replacements."""
    assert int(stats.population) == prev_population - int(stats.n_dead) + int(stats.n_born) + int(stats.n_replaced), \
        (prev_population, int(stats.n_dead), int(stats.n_born), int(stats.n_replaced), int(stats.population))
