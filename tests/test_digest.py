"""sugarscape_b200/digest.py: the order-independent state checksum bench.py prints as `state_digest` (equal at every N and
under every scheduler) and compares between the CUDA engine and the CPU oracle (`parity_at_scale`)."""
import numpy as np

import parity_util as pu
from sugarscape_b200 import digest, layout, sharded


def _state():
    c, state, pop, params, endow, tags = pu.build("combat_limited", layout.RNG_SCALABLE, {"agentMovement": [6, 6]})
    return state


def test_checksum_ignores_slots_and_list_order_but_not_the_agents():
    state = _state()
    base = digest.host_state_checksum(state)
    assert base["population"] == int(state.counters[layout.C_N_LIST])
    # the same simulation re-slotted for four ranks (other slots, other list order): same checksum
    new, lists = sharded.shard_host_state(state, 4)
    order = np.concatenate(lists[::-1])                       # (ranks in reverse: another list order again)
    new.a_order[:] = -1
    new.a_order[:len(order)] = order
    assert digest.host_state_checksum(new) == base
    # any change of an agent or of a cell changes it
    s = int(state.a_order[7])
    for field, delta in (("a_sugar", 1.0), ("a_age", 1), ("a_tribe", 1), ("a_pos", 1)):
        changed = state.copy()
        getattr(changed, field)[s] += delta
        assert digest.host_state_checksum(changed)["agents"] != base["agents"], field
    changed = state.copy()
    changed.sugar[123] += 1
    assert digest.host_state_checksum(changed)["grid"] != base["grid"]
    # two agents swapping their holdings is a different state, too (the agent's ID enters its term)
    a, b = int(state.a_order[3]), int(state.a_order[4])
    if state.a_sugar[a] != state.a_sugar[b]:
        changed = state.copy()
        changed.a_sugar[a], changed.a_sugar[b] = state.a_sugar[b], state.a_sugar[a]
        assert digest.host_state_checksum(changed)["agents"] != base["agents"]


def test_partial_checksums_of_the_slabs_add_up():
    state = _state()
    base = digest.host_state_checksum(state)
    new, lists = sharded.shard_host_state(state, 2)
    xb, H = new.slab_bounds, state.height
    parts = []
    for r, slots in enumerate(lists):
        lo, hi = xb[r] * H, xb[r + 1] * H
        parts.append({"population": len(slots),
                      "agents": digest.agents_checksum(new.a_id[slots], new.a_pos[slots], new.a_age[slots], new.a_sugar[slots],
                                                       new.a_spice[slots], new.a_tribe[slots]),
                      "grid": digest.grid_checksum(lo, new.sugar[lo:hi], new.spice[lo:hi], new.occ[lo:hi] >= 0)})
    assert digest.combine(parts) == base
    assert digest.as_text(base).split(":")[0] == str(base["population"])
