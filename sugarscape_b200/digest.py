"""Order-independent checksums of a simulation state (a checksum of checksums: one 64-bit word per agent / cell,
summed modulo 2^64).

Used where ID-sorted SHA digests of the whole state (tests/parity_util.py) are too expensive or impossible: by
``bench.py`` to show that the CPU oracle and the CUDA engine agree on the full-size S2 state ("parity_at_scale") and
that runs on 1 / 2 / 4 / 8 GPUs end in the same state ("state_digest" -- every rank sums its own agents and cells, the
partial sums add up).  The agent part identifies an agent by its ID, so slot numbers, list order and sharding do not
matter; the grid part covers sugar, spice and the occupied flag of every cell (which agent stands there is covered by
the agents' positions)."""
import numpy as np

_C = [np.uint64(v) for v in (0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB, 0xD6E8FEB86659FD93,
                             0xA0761D6478BD642F, 0xE7037ED1A0B428DB, 0x8EBC6AF09C88C6E3, 0x589965CC75374CC3)]


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * _C[1]
    z = (z ^ (z >> np.uint64(27))) * _C[2]
    return z ^ (z >> np.uint64(31))


def _u64(a):
    a = np.asarray(a)
    if a.dtype == np.float64:
        return a.view(np.uint64)
    return a.astype(np.int64).view(np.uint64)


def agents_checksum(ids, pos, age, sugar, spice, tribe):
    """Sum over the agents of mix(id, position, age, holdings (bit patterns), tribe)."""
    with np.errstate(over="ignore"):
        acc = _u64(ids) * _C[0]
        for k, field in enumerate((pos, age, sugar, spice, tribe)):
            acc = acc + _mix(_u64(field) + _C[3 + k]) * _C[(k + 1) % 3]
        return int(_mix(acc).sum(dtype=np.uint64))


def grid_checksum(first_cell, sugar, spice, occupied):
    """Sum over the cells [first_cell, first_cell + len) of (cell + 1) * (levels and occupied flag, weighted)."""
    with np.errstate(over="ignore"):
        n = len(sugar)
        cell = np.arange(first_cell + 1, first_cell + 1 + n, dtype=np.uint64)
        v = _u64(sugar) * _C[4] + _u64(spice) * _C[5] + np.asarray(occupied).astype(np.uint64) * _C[6] + _C[7]
        return int((cell * v).sum(dtype=np.uint64))


def host_state_checksum(state):
    """The checksums of a layout.HostState (the CPU oracle's state, or a downloaded engine state)."""
    from . import layout
    n = int(state.counters[layout.C_N_LIST])
    slots = state.a_order[:n]
    return {"population": n,
            "agents": agents_checksum(state.a_id[slots], state.a_pos[slots], state.a_age[slots], state.a_sugar[slots],
                                      state.a_spice[slots], state.a_tribe[slots]),
            "grid": grid_checksum(0, state.sugar, state.spice, state.occ >= 0)}


def engine_checksum(eng):
    """The same for a single-GPU Engine (downloads the eight arrays involved)."""
    hs = eng.download(["counters", "a_order", "a_id", "a_pos", "a_age", "a_sugar", "a_spice", "a_tribe", "sugar", "spice", "occ"])
    return host_state_checksum(hs)


def sharded_partial_checksum(eng):
    """This rank's part for a sharded engine: its own agent list and its own slab of the grid.  Sum the parts of all
    ranks modulo 2^64 (combine) for the checksum of the whole simulation."""
    from . import layout
    n = int(eng.counters()[layout.C_N_LIST])
    slots = eng.tensors["a_order"][:n].cpu().numpy().astype(np.int64)
    lo, hi = eng.slot_lo, eng.slot_hi
    inside = (slots >= lo) & (slots < hi)

    def field(name):
        local = eng.tensors[name][lo:hi].cpu().numpy()
        out = np.empty(n, dtype=local.dtype)
        out[inside] = local[slots[inside] - lo]
        for s in np.nonzero(~inside)[0]:                      # (agents in foreign slots: born across a slab boundary)
            out[s] = eng.tensors[name][int(slots[s]):int(slots[s]) + 1].cpu().numpy()[0]
        return out

    clo, chi = eng.cell_lo, eng.cell_hi
    return {"population": n,
            "agents": agents_checksum(field("a_id"), field("a_pos"), field("a_age"), field("a_sugar"), field("a_spice"), field("a_tribe")),
            "grid": grid_checksum(clo, eng.tensors["sugar"][clo:chi].cpu().numpy(), eng.tensors["spice"][clo:chi].cpu().numpy(),
                                  eng.tensors["occ"][clo:chi].cpu().numpy() >= 0)}


def combine(parts):
    mask = (1 << 64) - 1
    return {"population": sum(p["population"] for p in parts),
            "agents": sum(p["agents"] for p in parts) & mask, "grid": sum(p["grid"] for p in parts) & mask}


def as_text(chk):
    return f"{chk['population']}:{chk['agents']:016x}:{chk['grid']:016x}"
