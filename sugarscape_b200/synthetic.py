"""Synthetic large-map workloads (BASELINE.json configs S1 / S2).

The reference cannot instantiate these shapes (about 3.4 KB of Python objects per cell), so the
state is generated directly as struct-of-arrays with numpy, following the reference's endowment
scheme: every ranged endowment is the cyclic sequence min..max (reference sugarscape.py:697-703)
in a random permutation (725-729), sexes are split by agentMaleToFemaleRatio (692-717) and agents
are placed without replacement on cells of the active quadrants.  The capacity field is the
periodic tiling cap(x, y) = cap50(x mod 50, y mod 50) of the reference's default 50 x 50
two-peak field, because the shipped two-peak formula leaves > 99 % of a 4096^2 map barren.
Runs in scalable RNG mode only.
"""
import numpy as np

from . import config, init, layout

S1_OPTIONS = {   # examples/reproduction_basic.json rules on a 4096^2 map with 1.6 M agents
    "agentFemaleInfertilityAge": [40, 50], "agentFemaleFertilityAge": [12, 15], "agentFertilityFactor": [1, 1],
    "agentMaleInfertilityAge": [50, 60], "agentMaleFertilityAge": [12, 15], "agentMaxAge": [60, 100],
    "agentMovement": [6, 6], "agentStartingSugar": [50, 100], "seed": 12345,
    "environmentWidth": 4096, "environmentHeight": 4096, "startingAgents": 1600000, "timesteps": 100,
}
S2_OPTIONS = {   # examples/combat_limited.json rules on a 16384^2 map with 25 M agents
    "agentMaxAge": [60, 100], "agentMovement": [6, 6], "agentStartingSugar": [50, 100],
    "agentTagStringLength": 11, "environmentMaxCombatLoot": 2, "environmentMaxTribes": 2,
    "environmentQuadrantSizeFactor": 0.8, "environmentStartingQuadrants": [2, 4],
    "environmentTribePerQuadrant": True, "seed": 12345,
    "environmentWidth": 16384, "environmentHeight": 16384, "startingAgents": 25000000,
    "agentReplacements": 25000000, "timesteps": 100,
}


def configuration(options, overrides=None):
    c = config.overlay_options(dict(options), config.default_configuration())
    c.update(overrides or {})
    c["headlessMode"] = True
    c["b200RngMode"] = "scalable"
    # no verify_random_seed here: synthetic states do not use the CPython stream
    return config.verify_configuration(c)


def _cyclic(rng, lo, hi, n, dtype):
    values = lo + (np.arange(n, dtype=np.int64) % (hi - lo + 1))
    return rng.permutation(values).astype(dtype)


def tiled_capacity(width, height, peaks, tile=50):
    base = init.capacity_field(tile, tile, peaks)
    xs = np.arange(width) % tile
    ys = np.arange(height) % tile
    return base[np.ix_(xs, ys)].astype(np.int32)


def build_state(c, capacity=None, seed=None):
    """Returns (HostState, params)."""
    W, H, n = c["environmentWidth"], c["environmentHeight"], int(c["startingAgents"])
    rng = np.random.default_rng(c["seed"] if seed is None else seed)
    if capacity is None:
        capacity = int(min(max(3 * n, 1024), 2 * W * H + 1024)) if c["agentFertilityFactor"][1] > 0 else int(n + n // 2 + 1024)
    state = layout.HostState.empty(W, H, capacity, kin_capacity=16)     # kin lists are a mode E feature
    state.sugar_cap[:] = tiled_capacity(W, H, c["environmentSugarPeaks"]).reshape(-1)
    state.sugar[:] = state.sugar_cap
    if c["environmentMaxSpice"] > 0:
        state.spice_cap[:] = tiled_capacity(W, H, c["environmentSpicePeaks"]).reshape(-1)
        state.spice[:] = state.spice_cap
    # placement: uniform without replacement over the active quadrants
    quadrants = c["environmentStartingQuadrants"]
    qw = int(np.floor(W / 2 * c["environmentQuadrantSizeFactor"]))
    qh = int(np.floor(H / 2 * c["environmentQuadrantSizeFactor"]))
    spans = {1: (0, 0), 2: (W - qw, 0), 3: (W - qw, H - qh), 4: (0, H - qh)}
    per_q = [n // len(quadrants) + (1 if i < n % len(quadrants) else 0) for i in range(len(quadrants))]
    cells, tribe_of = [], []
    for qi, (q, k) in enumerate(zip(quadrants, per_q)):
        x0, y0 = spans[q]
        local = rng.choice(qw * qh, size=k, replace=False)
        cells.append(((x0 + local // qh) * H + (y0 + local % qh)).astype(np.int64))
        tribe_of.append(np.full(k, qi, dtype=np.int32))
    cells = np.concatenate(cells)
    tribe_of = np.concatenate(tribe_of)
    perm = rng.permutation(n)
    cells, tribe_of = cells[perm], tribe_of[perm]
    sl = slice(0, n)
    state.a_id[sl] = np.arange(n, dtype=np.int32)
    state.a_pos[sl] = cells
    state.a_order[sl] = np.arange(n, dtype=np.int32)
    state.occ[cells] = np.arange(n, dtype=np.int32)
    sugar = _cyclic(rng, *c["agentStartingSugar"], n, np.float64)
    state.a_sugar[sl] = sugar
    state.a_start_sugar[sl] = sugar
    if c["agentStartingSpice"][1] > 0:
        spice = _cyclic(rng, *c["agentStartingSpice"], n, np.float64)
        state.a_spice[sl] = spice
        state.a_start_spice[sl] = spice
    state.a_sugar_metab[sl] = _cyclic(rng, *c["agentSugarMetabolism"], n, np.int32)
    state.a_spice_metab[sl] = _cyclic(rng, *c["agentSpiceMetabolism"], n, np.int32)
    state.a_vision[sl] = _cyclic(rng, *c["agentVision"], n, np.int32)
    state.a_movement[sl] = _cyclic(rng, *c["agentMovement"], n, np.int32)
    state.a_max_age[sl] = _cyclic(rng, *c["agentMaxAge"], n, np.int32) if c["agentMaxAge"][0] >= 0 else -1
    ratio = c["agentMaleToFemaleRatio"]
    if ratio:
        males = int(np.floor(n / (ratio + 1)) * ratio)
        sex = np.where(rng.permutation(n) < males, layout.SEX_MALE, layout.SEX_FEMALE).astype(np.int32)
        state.a_sex[sl] = sex
        f_fa = _cyclic(rng, *c["agentFemaleFertilityAge"], n, np.int32)
        f_ia = _cyclic(rng, *c["agentFemaleInfertilityAge"], n, np.int32)
        m_fa = _cyclic(rng, *c["agentMaleFertilityAge"], n, np.int32)
        m_ia = _cyclic(rng, *c["agentMaleInfertilityAge"], n, np.int32)
        state.a_fert_age[sl] = np.where(sex == layout.SEX_FEMALE, f_fa, m_fa)
        state.a_infert_age[sl] = np.where(sex == layout.SEX_FEMALE, f_ia, m_ia)
    state.a_fert_factor[sl] = _cyclic(rng, *c["agentFertilityFactor"], n, np.float64)
    state.a_aggression[sl] = _cyclic(rng, *c["agentAggressionFactor"], n, np.float64)
    if c["agentTagStringLength"] > 0 and c["environmentMaxTribes"] > 0:
        L, T = c["agentTagStringLength"], c["environmentMaxTribes"]
        tribe = tribe_of if c["environmentTribePerQuadrant"] else (np.arange(n) % T).astype(np.int32)
        size = (L + 1) / T
        lo = np.floor(tribe * size).astype(np.int64)
        hi = np.minimum(np.floor((tribe + 1) * size).astype(np.int64) - 1, L)
        zeroes = lo + (rng.integers(0, 1 << 30, n) % (hi - lo + 1))
        shifts = np.arange(L, dtype=np.uint32)[None, :]
        for b0 in range(0, n, 1 << 21):                      # chunked: n x L temporaries are large at S2
            b1 = min(n, b0 + (1 << 21))
            ranks = rng.random((b1 - b0, L)).argsort(axis=1).argsort(axis=1)
            bits = (ranks >= zeroes[b0:b1, None]).astype(np.uint32)     # `zeroes` zero bits at random positions
            state.a_tags[b0:b1] = (bits << shifts).sum(axis=1).astype(np.uint32)
            z = L - bits.sum(axis=1)
            state.a_tribe[b0:b1] = np.minimum(np.ceil((z + 1) / size).astype(np.int64) - 1, T - 1)
    state.counters[layout.C_N_LIST] = n
    state.counters[layout.C_N_SLOTS] = n
    state.counters[layout.C_NEXT_ID] = n
    params = layout.make_params(c, layout.RNG_SCALABLE, init.rule_flags(c), init.max_agent_range(c), init.equator(c))
    return state, params


def endowment_table(state, n):
    """The endowment list replacements cycle through (reference sugarscape.py:171-267 pops the
    configured endowments in creation order): here the starting agents' own endowments."""
    src = {"sugar": "a_start_sugar", "spice": "a_start_spice", "sugar_metab": "a_sugar_metab",
           "spice_metab": "a_spice_metab", "vision": "a_vision", "movement": "a_movement", "max_age": "a_max_age",
           "sex": "a_sex", "fert_age": "a_fert_age", "infert_age": "a_infert_age", "fert_factor": "a_fert_factor",
           "aggression": "a_aggression", "lookahead": "a_lookahead", "trade_factor": "a_trade_factor", "tags": "a_tags"}
    return {name: np.ascontiguousarray(getattr(state, src[name])[:n]).astype(dt) for name, dt in layout.ENDOWMENT_FIELDS}
