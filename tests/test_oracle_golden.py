"""The CPU oracle (oracle/ss_oracle.c) against fixtures generated from the UNMODIFIED reference
(tests/golden/make_golden.py): every shipped example, every step of the reference's
200-step test protocol -- agents (IDs, positions, ages in list order), fp64 holdings, tags, grid,
pollution and the CPython MT19937 state, all bit-exact; the 59 log keys json-equal.  The 24 examples
the engine supports run through the ABI state alone; the other five (lending, friends, diseases,
depression, racial tags, universal income, ethics decision models) through oracle-private state,
ahead of the engine, which still refuses them."""
import math

import pytest

import parity_util as pu
from oracle_binding import Oracle
from sugarscape_b200 import init, stats

SUPPORTED = [
    "agent_replacement", "blank", "combat_limited", "combat_unlimited", "constant_growback",
    "cultural_combat_unlimited", "cultural_tagging", "diagonal_waves", "foresight_basic",
    "immediate_growback", "pollution_formation", "reproduction_basic", "reproduction_inheritance",
    "reproduction_oscillation_large", "reproduction_oscillation_severe",
    "reproduction_oscillation_small", "seasonal_migration", "spice_growback", "trade_basic",
    "trade_pollution", "trade_replacement", "trade_reproduction", "trade_tagging",
    "trade_tagging_reproduction",
]
ALL_FEATURES = ["all_features", "determinism", "disease_basic", "dune", "lending_basic"]      # beyond the ABI 1 state


def _close(a, b):
    if a is None or b is None:
        return a is b
    if isinstance(a, float) or isinstance(b, float):
        return math.isclose(a, b, rel_tol=1e-9, abs_tol=1e-9)
    return a == b


@pytest.mark.parametrize("name", SUPPORTED)
def test_oracle_matches_reference_trajectory(name):
    c, state, pop, params, endow, tags = pu.build(name)
    meta, _ = pu.load_golden(name)
    gold, log = meta["steps"], meta["log"]
    orc = Oracle(params, state, endow, tags)
    orc.set_mt_from_python()
    d, _ = pu.state_digests(state, *orc.mt_state())
    assert d == gold[0]["digests"], "t=0 state differs"
    sb = stats.StatsBuilder(c, init.equator(c))
    rs = sb.update(orc.stats(0), 0)
    skip = set()      # all 59 keys are checked
    for t in range(1, len(gold)):
        orc.env_step(t)
        orc.agent_step(t, rs["meanWealth"])
        orc.compact(t)
        replaced = 0
        if c["agentReplacements"] > 0:
            orc.push_mt_to_python()
            replaced = pu.replace_dead_agents(c, state, pop, t)
            orc.set_mt_from_python()
        d, snap = pu.state_digests(state, *orc.mt_state())
        assert d == gold[t]["digests"], f"state differs at t={t}"
        assert len(snap["a_id"]) == gold[t]["population"]
        rs = sb.update(orc.stats(t), t, replaced)
        for k in stats.LOG_KEYS:
            if k not in skip:
                assert _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs[k]} vs {log[t][k]}"


def test_oracle_lending_matches_reference_trajectory():
    """Lending (reference agent.py:431-483 and helpers) exists in the ORACLE only so far -- the engine
    refuses lending configurations.  The restatement is pinned here against the unmodified
    reference on examples/lending_basic.json: every digest of every step, the loanVolume log key
    and the other 58 keys."""
    name = "lending_basic"
    c, state, pop, params, endow, tags = pu.build(name, check=False)
    meta, _ = pu.load_golden(name)
    gold, log = meta["steps"], meta["log"]
    orc = Oracle(params, state, endow, tags)
    orc.set_mt_from_python()
    orc.bind_lending(pop.endowments)
    try:
        d, _ = pu.state_digests(state, *orc.mt_state())
        assert d == gold[0]["digests"], "t=0 state differs"
        sb = stats.StatsBuilder(c, init.equator(c))
        rs = sb.update(orc.stats(0), 0)
        volume = 0
        for t in range(1, len(gold)):
            orc.env_step(t)
            orc.agent_step(t, rs["meanWealth"])
            orc.compact(t)
            d, snap = pu.state_digests(state, *orc.mt_state())
            assert d == gold[t]["digests"], f"state differs at t={t}"
            rs = sb.update(orc.stats(t), t, 0)
            assert orc.loan_volume(t) == log[t]["loanVolume"], f"loanVolume differs at t={t}"
            volume += orc.loan_volume(t)
            for k in stats.LOG_KEYS:
                if k != "loanVolume":
                    assert _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs[k]} vs {log[t][k]}"
        assert volume > 1000          # the rule was exercised (about 2.5 k loans in 200 steps)
    finally:
        orc.unbind_lending()


def test_oracle_dune_lending_friends_combat_matches_reference_trajectory():
    """examples/dune.json = trade + tags + live combat + reproduction + inheritance to children on top
    of lending (payDebtToCreditorChildren included) and friends.  Oracle-only restatements (the
    engine refuses the configuration): every digest of every step and all 59 log keys, with the
    social / conflict happiness means and loanVolume taken from the oracle-private state."""
    name = "dune"
    c, state, pop, params, endow, tags = pu.build(name, check=False)
    meta, _ = pu.load_golden(name)
    gold, log = meta["steps"], meta["log"]
    orc = Oracle(params, state, endow, tags)
    orc.set_mt_from_python()
    orc.bind_lending(pop.endowments)
    orc.bind_social(pop.endowments)
    try:
        d, _ = pu.state_digests(state, *orc.mt_state())
        assert d == gold[0]["digests"], "t=0 state differs"
        sb = stats.StatsBuilder(c, init.equator(c))
        rs = sb.update(orc.stats(0), 0)
        fights = 0
        for t in range(1, len(gold)):
            orc.env_step(t)
            orc.agent_step(t, rs["meanWealth"])
            orc.compact(t)
            d, snap = pu.state_digests(state, *orc.mt_state())
            assert d == gold[t]["digests"], f"state differs at t={t}"
            rs = sb.update(orc.stats(t), t, 0)
            n = rs["population"]
            social, conflict = orc.social_sums()
            rs["meanSocialHappiness"] = round(social / n, 2) if n else 0
            rs["meanConflictHappiness"] = round(conflict / n, 2) if n else 0
            rs["loanVolume"] = orc.loan_volume(t)
            fights += conflict != 0
            for k in stats.LOG_KEYS:
                assert _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs[k]} vs {log[t][k]}"
        assert fights > 50      # live combat moved the conflict happiness in most early steps
    finally:
        orc.unbind_lending()
        orc.unbind_social()


def _disease_case(base, overrides, gold, log):
    c, state, pop, params, endow, tags = pu.build(base, overrides=overrides, check=False)
    orc = Oracle(params, state, endow, tags)
    orc.bind_disease(c, pop.endowments, pop.disease_endowments, len(gold) + 8)     # initial infections use the main stream
    orc.set_mt_from_python()
    try:
        d, _ = pu.state_digests(state, *orc.mt_state())
        assert d == gold[0]["digests"], "t=0 state differs"
        sb = stats.StatsBuilder(c, init.equator(c))
        rs = sb.update(orc.stats(0), 0)
        rs.update(orc.disease_log_keys(0, rs["population"]))
        for k in ("sickAgents", "diseaseIncidence", "diseasePrevalence"):
            assert rs[k] == log[0][k], f"{k} differs after init"
        sick_steps = 0
        for t in range(1, len(gold)):
            orc.env_step(t)
            orc.agent_step(t, rs["meanWealth"])
            orc.compact(t)
            d, snap = pu.state_digests(state, *orc.mt_state())
            assert d == gold[t]["digests"], f"state differs at t={t}"
            rs = sb.update(orc.stats(t), t, 0)
            rs.update(orc.disease_log_keys(t, rs["population"]))
            sick_steps += rs["sickAgents"] > 0
            for k in stats.LOG_KEYS:
                assert _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs[k]} vs {log[t][k]}"
        return sick_steps
    finally:
        orc.unbind_disease()


def test_oracle_disease_basic_matches_reference_trajectory():
    """examples/disease_basic.json (oracle-only restatement of diseases; the engine refuses them):
    the population is extinct after five steps, every digest and all 59 log keys on the way."""
    meta, _ = pu.load_golden("disease_basic")
    assert _disease_case("disease_basic", None, meta["steps"], meta["log"]) >= 4


@pytest.mark.parametrize("variant,min_sick_steps", [("disease_mild", 15), ("disease_endemic", 80), ("disease_late_start", 25)])
def test_oracle_disease_variants_match_reference_trajectory(variant, min_sick_steps):
    """disease_basic with milder diseases (tests/golden/make_golden_variants.py, run on the unmodified
    reference): incubation, transmission / protection draws, recovery through the immune response,
    births with inherited immune systems, vision pushed to 0 -- digests + 59 log keys on every step."""
    import gzip
    import json
    import os
    with gzip.open(os.path.join(pu.GOLDEN, variant + ".json.gz"), "rt") as f:
        meta = json.load(f)
    overrides = dict(meta["overrides"], timesteps=meta["meta"]["timesteps"])
    assert _disease_case(meta["base"], overrides, meta["steps"], meta["log"]) >= min_sick_steps


def _bind_everything(orc, c, pop, horizon, ethics):
    orc.bind_lending(pop.endowments)
    orc.bind_social(pop.endowments)
    orc.bind_misc(c, pop.endowments, horizon)
    if ethics:
        orc.bind_ethics(c, pop.endowments)
    orc.bind_disease(c, pop.endowments, pop.disease_endowments, horizon)      # (initial infections draw from the main stream)
    orc.set_mt_from_python()


def _unbind_everything(orc):
    orc.unbind_lending()
    orc.unbind_social()
    orc.unbind_misc()
    orc.unbind_ethics()
    orc.unbind_disease()


def _private_log_keys(orc, rs, t, ethics):
    """The log keys that come from the oracle-private rule state."""
    n = rs["population"]
    rs.update(orc.disease_log_keys(t, n))
    social, conflict = orc.social_sums()
    rs["meanSocialHappiness"] = round(social / n, 2) if n else 0
    rs["meanConflictHappiness"] = round(conflict / n, 2) if n else 0
    rs["loanVolume"] = orc.loan_volume(t)
    if n:
        rs["largestRace"], rs["largestRaceSize"], rs["remainingRaces"] = orc.race_census()
    if ethics:
        rs.update(orc.ethics_log_keys(t, n, rs["meanNeighbors"]))


@pytest.mark.parametrize("variant,ethics", [("determinism_plain", False), ("determinism_ethics", True), ("nowrap_determinism", True), ("moore_nowrap_determinism", True),
                                            ("radial_determinism", True), ("radial_nowrap_determinism", True), ("determinism_late_diseases", True)])
def test_oracle_determinism_variants_match_reference_trajectory(variant, ethics):
    """examples/determinism.json (variant fixtures from the unmodified reference,
    tests/golden/make_golden_variants.py): diseases, depression, lending, friends, racial tags,
    universal income, live combat, trade, tags, inheritance, pollution and seasons at once --
    `_plain` without, `_ethics` with the ethics.Bentham decision models (bentham, altruist, egoist,
    negativeBentham: SURVEY row a21; about a quarter of the moves are not the greedy one).  Only the
    experimental-group split of the log is switched off.  Every digest and all 59 log keys for 200
    steps.  Oracle-only: the engine refuses these rule families."""
    import gzip
    import json
    import os
    with gzip.open(os.path.join(pu.GOLDEN, variant + ".json.gz"), "rt") as f:
        meta = json.load(f)
    gold, log = meta["steps"], meta["log"]
    overrides = dict(meta["overrides"], timesteps=meta["meta"]["timesteps"])
    c, state, pop, params, endow, tags = pu.build(meta["base"], overrides=overrides, check=False)
    orc = Oracle(params, state, endow, tags)
    _bind_everything(orc, c, pop, len(gold) + 8, ethics)
    try:
        d, _ = pu.state_digests(state, *orc.mt_state())
        assert d == gold[0]["digests"], "t=0 state differs"
        sb = stats.StatsBuilder(c, init.equator(c))
        rs = sb.update(orc.stats(0), 0)
        non_greedy = 0
        for t in range(1, len(gold)):
            orc.env_step(t)
            orc.agent_step(t, rs["meanWealth"])
            orc.compact(t)
            d, snap = pu.state_digests(state, *orc.mt_state())
            assert d == gold[t]["digests"], f"state differs at t={t}"
            rs = sb.update(orc.stats(t), t, 0)
            _private_log_keys(orc, rs, t, ethics)
            non_greedy += rs["agentLastMoveOptimalityPercentage"] < 100
            for k in stats.LOG_KEYS:
                assert _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs[k]} vs {log[t][k]}"
        assert (non_greedy > 0.7 * (len(gold) - 1)) == ethics
    finally:
        _unbind_everything(orc)


@pytest.mark.parametrize("variant", ["ethics_trade", "ethics_combat", "ethics_replacement", "ethics_pollution"])
def test_oracle_ethics_variants_match_reference_trajectory(variant):
    """The ethics.Bentham restatement alone (no lending / diseases / friends around it), on rule families the
    engine supports: fixtures from the unmodified reference (make_golden_variants.py) with bentham, altruist,
    egoist, negativeBentham and the NoLookahead / HalfLookahead spawn variants over trade + tagging +
    reproduction, live combat, host-side replacement and pollution.  Digests and all 59 log keys."""
    from sugarscape_b200 import layout
    base, overrides, gold, log = pu.load_variant(variant)
    c, state, pop, params, endow, tags = pu.build(base, overrides=overrides)
    orc = Oracle(params, state, endow, tags)
    orc.bind_ethics(c, pop.endowments)
    orc.set_mt_from_python()
    try:
        d, _ = pu.state_digests(state, *orc.mt_state())
        assert d == gold[0]["digests"], "t=0 state differs"
        sb = stats.StatsBuilder(c, init.equator(c))
        rs = sb.update(orc.stats(0), 0, 0, orc.ethics_stats(0))
        for t in range(1, len(gold)):
            orc.env_step(t)
            orc.agent_step(t, rs["meanWealth"])
            orc.compact(t)
            replaced = 0
            n = int(state.counters[layout.C_N_LIST])
            if 0 < c["agentReplacements"] and n < c["agentReplacements"]:
                orc.push_mt_to_python()
                new_agents = pop.place(c["agentReplacements"] - n, state.occ >= 0, t, int(state.counters[layout.C_NEXT_ID]))
                state.append_agents(new_agents, protect_last=int(state.counters[layout.C_N_DEAD]))
                state.counters[layout.C_NEXT_ID] += len(new_agents)
                orc.adopt_new_agents(new_agents, c)
                orc.set_mt_from_python()
                replaced = len(new_agents)
            d, snap = pu.state_digests(state, *orc.mt_state())
            assert d == gold[t]["digests"], f"state differs at t={t}"
            rs = sb.update(orc.stats(t), t, replaced, orc.ethics_stats(t))
            for k in stats.LOG_KEYS:
                assert _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs[k]} vs {log[t][k]}"
    finally:
        orc.unbind_ethics()


_INTERACTIONS = ("combat", "disease", "lending", "reproduction", "trade")
# keys a group pass does not compute: the log keeps their initial 0 (sugarscape.py:1403-1417)
_GROUP_TEMPLATE_ZERO = ("timestep", "giniCoefficient", "seed", "environmentWealthCreated", "environmentWealthTotal")


def _full_log_entry(orc, sb, group, t):
    """The reference's runtimeStats with an experimental group (sugarscape.py:998-1003): the statistics of
    the group, of its complement ("control") and of everybody -- 59 + 2 + 2 x (61 + 10) = 203 keys."""
    import copy

    def one_pass(which, builder):
        orc.set_pass(which)
        rs = builder.update(orc.stats(t), t, 0)
        _private_log_keys(orc, rs, t, True)
        counters = orc.group_counters()
        n = rs["population"]
        rs["meanControlNeighbors"] = round(counters[10] / n, 2) if n else 0
        rs["meanExperimentalNeighbors"] = round(counters[11] / n, 2) if n else 0
        return rs, counters

    keys = tuple(stats.LOG_KEYS) + ("meanControlNeighbors", "meanExperimentalNeighbors")
    passes = {group: one_pass(1, copy.deepcopy(sb)), "control": one_pass(2, copy.deepcopy(sb))}
    everybody, _ = one_pass(0, sb)
    entry = {k: everybody[k] for k in keys}
    for prefix, (rs, counters) in passes.items():
        rs["carryingCapacity"] = entry["carryingCapacity"]          # from len(self.agents), whatever the pass
        for k in keys:
            entry[prefix + k[0].upper() + k[1:]] = 0 if k in _GROUP_TEMPLATE_ZERO else rs[k]
        side = "Control" if prefix == "control" else "Experimental"
        for i, kind in enumerate(_INTERACTIONS):
            entry[f"{kind}{side}GroupToControlGroup"] = counters[i]
            entry[f"{kind}{side}GroupToExperimentalGroup"] = counters[5 + i]
    return entry


@pytest.mark.parametrize("name", ["determinism", "all_features"])
def test_oracle_all_features_examples_match_reference_trajectory(name):
    """examples/determinism.json (the example BASELINE.json's north star names) and all_features.json as
    shipped, oracle only: every digest of every step and ALL 203 log keys -- the 59 base keys, the
    movement-neighbourhood split and the complete experimental-group ("depressed") / control copies
    with their interaction counters."""
    c, state, pop, params, endow, tags = pu.build(name, check=False)
    meta, _ = pu.load_golden(name)
    gold, log = meta["steps"], meta["log"]
    orc = Oracle(params, state, endow, tags)
    orc.bind_group()
    _bind_everything(orc, c, pop, len(gold) + 8, True)
    try:
        d, _ = pu.state_digests(state, *orc.mt_state())
        assert d == gold[0]["digests"], "t=0 state differs"
        sb = stats.StatsBuilder(c, init.equator(c))
        rs = _full_log_entry(orc, sb, c["experimentalGroup"], 0)
        assert len(log[1]) == 203
        for t in range(0, len(gold)):
            if t > 0:
                orc.env_step(t)
                orc.agent_step(t, rs["meanWealth"])
                orc.compact(t)
                d, snap = pu.state_digests(state, *orc.mt_state())
                assert d == gold[t]["digests"], f"state differs at t={t}"
                rs = _full_log_entry(orc, sb, c["experimentalGroup"], t)
            for k in log[t]:
                assert k in rs and _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs.get(k)} vs {log[t][k]}"
    finally:
        orc.unbind_group()
        _unbind_everything(orc)


def test_oracle_environment_file_matches_reference_trajectory():
    """`environmentFile` (reference sugarscape.py:370-384, incl. its sugar / spice swap): trade + tagging +
    reproduction rules on the 30 x 30 map of tests/golden/environment_30.json, variant fixture from the
    unmodified reference.  The engine supports this option (it only changes the initial arrays)."""
    base, overrides, gold, log = pu.load_variant("environment_file")
    c, state, pop, params, endow, tags = pu.build(base, overrides=overrides)
    orc = Oracle(params, state, endow, tags)
    orc.set_mt_from_python()
    d, _ = pu.state_digests(state, *orc.mt_state())
    assert d == gold[0]["digests"], "t=0 state differs"
    sb = stats.StatsBuilder(c, init.equator(c))
    rs = sb.update(orc.stats(0), 0)
    for t in range(1, len(gold)):
        orc.env_step(t)
        orc.agent_step(t, rs["meanWealth"])
        orc.compact(t)
        d, snap = pu.state_digests(state, *orc.mt_state())
        assert d == gold[t]["digests"], f"state differs at t={t}"
        rs = sb.update(orc.stats(t), t, 0)
        for k in stats.LOG_KEYS:
            assert _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs[k]} vs {log[t][k]}"


@pytest.mark.parametrize("overrides", [{"agentVisionMode": "radial", "agentMovementMode": "radial", "b200RngMode": "scalable"},
                                       {"agentVisionMode": "radial", "agentMovementMode": "radial", "environmentWraparound": False, "b200RngMode": "scalable"},
                                       {"agentVisionMode": "radial", "agentMovementMode": "radial", "agentVision": [1, 8], "agentMovement": [8, 8]},
                                       {"neighborhoodMode": "moore", "b200RngMode": "scalable"},
                                       {"environmentWraparound": False, "b200RngMode": "scalable"},
                                       {"environmentWraparound": False, "agentVision": [1, 30], "agentMovement": [30, 30]},
                                       {"agentDecisionModels": ["asimov", "negativeBentham"]}, {"agentDecisionModels": ["asimovTop"]},
                                       {"agentDecisionModels": ["negativeBenthamTop"]}, {"experimentalGroup": "raceInGroup"}, {"experimentalGroup": "disease3"}, {"experimentalGroup": "ageRange0", "environmentAgeistAbsoluteRanges": [[0, 30]]},
                                       {"agentLeader": True}, {"agentMaxFriends": [20, 20]}])
def test_unsupported_options_fail_loudly(overrides):
    """All 29 shipped examples are supported; the options outside the CUDA implementation are refused, not half-run."""
    with pytest.raises(init.UnsupportedConfiguration):
        pu.build("determinism", overrides=overrides)


@pytest.mark.parametrize("name,mode", [("reproduction_basic", "exact"), ("trade_tagging_reproduction", "exact"),
                                       ("combat_limited", "scalable"), ("agent_replacement", "scalable"),
                                       ("reproduction_basic", "scalable"), ("trade_replacement", "scalable")])
def test_state_invariants_hold_on_the_oracle(name, mode):
    """tests/invariants.py (the checks the GPU tests apply at BASELINE's full sizes) on oracle states."""
    import invariants
    from sugarscape_b200 import layout
    rng_mode = layout.RNG_EXACT if mode == "exact" else layout.RNG_SCALABLE
    c, state, pop, params, endow, tags = pu.build(name, rng_mode, {"agentMovement": [6, 6]})
    orc = Oracle(params, state, endow, tags)
    if rng_mode == layout.RNG_EXACT:
        orc.set_mt_from_python()
    elif c["agentReplacements"] > 0:
        orc.bind_endowments(pop.endowments)
    population = invariants.check_state(state)
    for t in range(1, 41):
        orc.env_step(t)
        orc.agent_step(t, 0.0)
        orc.compact(t)
        if c["agentReplacements"] > 0:
            if rng_mode == layout.RNG_EXACT:
                orc.push_mt_to_python()
                pu.replace_dead_agents(c, state, pop, t)
                orc.set_mt_from_python()
            else:
                orc.replace(t)
        st = orc.stats(t)
        if rng_mode == layout.RNG_SCALABLE:
            invariants.check_step_accounting(population, st)
        population = invariants.check_state(state, st)


def _load_private_scalable():
    import gzip
    import json
    import os
    with gzip.open(os.path.join(pu.GOLDEN, "scalable_mode_private.json.gz"), "rt") as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["determinism", "disease_basic", "dune", "lending_basic"])
def test_scalable_mode_private_families_match_patched_reference(name):
    """The scalable-mode definitions extended to the oracle-only rule families: lending draws from site 5,
    everything a disease step draws (shuffles, picks, infection attempts) from site 6 of the acting agent.
    Fixtures: the unmodified reference with its randomness redirected (make_golden_scalable.py --private);
    inheritance is off in these cases because the scalable mode keeps no kin lists."""
    from sugarscape_b200 import layout
    case = _load_private_scalable()[name]
    c, state, pop, params, endow, tags = pu.build(name, layout.RNG_SCALABLE, case["overrides"], check=False)
    orc = Oracle(params, state, endow, tags)
    horizon = len(case["steps"]) + 8
    orc.bind_lending(pop.endowments)
    orc.bind_social(pop.endowments)
    orc.bind_misc(c, pop.endowments, horizon)
    orc.bind_ethics(c, pop.endowments)
    diseases = c["agentImmuneSystemLength"] > 0
    if diseases:
        orc.bind_disease(c, pop.endowments, pop.disease_endowments, horizon)
    try:
        for rec in case["steps"]:
            t = rec["t"]
            orc.env_step(t)
            orc.agent_step(t, 0.0)
            orc.compact(t)
            d, snap = pu.state_digests_by_id(state)
            assert len(snap["a_id"]) == rec["population"], f"population differs at t={t}"
            for k in ("agents_int", "agents_f64", "agents_tags", "grid_int", "pollution"):
                assert d[k] == rec["digests"][k], f"{k} differs from the patched reference at t={t}"
    finally:
        _unbind_everything(orc)


def test_scalable_mode_priority_is_a_bijection():
    from oracle_binding import lib
    L = lib()
    key = L.orc_step_key(12345, 7)
    seen = {L.orc_priority(key, i, 12) for i in range(1 << 12)}
    assert len(seen) == 1 << 12


SCALABLE_GOLDEN = pu.load_scalable_golden()


@pytest.mark.parametrize("name", sorted(SCALABLE_GOLDEN))
def test_scalable_mode_oracle_matches_patched_reference(name):
    """Mode S is the reference's rule code with only its randomness redirected (priority order,
    per-agent per-site substreams, newborn numbered by cell): tests/golden/make_golden_scalable.py
    runs the unmodified reference classes under exactly those substitutions.  The oracle in mode S
    must reproduce every step: integer state bit-exact, fp64 holdings and pollution bit-exact too
    (same libm)."""
    from sugarscape_b200 import layout
    case = SCALABLE_GOLDEN[name]
    c, state, pop, params, endow, tags = pu.build(name, layout.RNG_SCALABLE, case["overrides"])
    orc = Oracle(params, state, endow, tags)
    orc.bind_endowments(pop.endowments)
    for rec in case["steps"]:
        t = rec["t"]
        orc.env_step(t)
        orc.agent_step(t, 0.0)
        orc.compact(t)
        orc.replace(t)
        d, snap = pu.state_digests_by_id(state)
        assert len(snap["a_id"]) == rec["population"], f"population differs at t={t}"
        for k in ("agents_int", "agents_f64", "agents_tags", "grid_int", "pollution"):
            assert d[k] == rec["digests"][k], f"{k} differs from the patched reference at t={t}"


# nowrap_*: environmentWraparound = false (cell.py:47-121) -- the neighbour lists of tagging / trade / births and the pollution
# flux lose the cells across the map's edge (with the reference's own south-neighbour test)
PLAIN_VARIANTS = ["pollution_diffusion_delay2", "pollution_diffusion_delay3", "nowrap_trade_tagging_reproduction", "nowrap_pollution",
                  "moore_trade_tagging_reproduction", "moore_pollution",       # moore_*: neighborhoodMode "moore" (cell.py:77-89)
                  # radial_*: agentVisionMode = agentMovementMode = "radial" (environment.py:146-173) -- discs instead of crosses
                  "radial_trade_tagging_reproduction", "radial_combat", "radial_small_map", "radial_nowrap_small_map"]


@pytest.mark.parametrize("variant", PLAIN_VARIANTS)
def test_oracle_plain_variants_match_reference_trajectory(variant):
    """Shipped examples with option overrides that only touch rule families of ABI 1 (fixtures from the unmodified
    reference, tests/golden/make_golden_variants.py).  pollution_diffusion_delay*: a diffusion timeframe that is
    open from the first step with a delay > 1 -- the countdown of environment.py:192-197 fires on t = 2, 4, ..."""
    base, overrides, gold, log = pu.load_variant(variant)
    c, state, pop, params, endow, tags = pu.build(base, overrides=overrides)
    orc = Oracle(params, state, endow, tags)
    orc.set_mt_from_python()
    d, _ = pu.state_digests(state, *orc.mt_state())
    assert d == gold[0]["digests"], "t=0 state differs"
    sb = stats.StatsBuilder(c, init.equator(c))
    rs = sb.update(orc.stats(0), 0)
    for t in range(1, len(gold)):
        orc.env_step(t)
        orc.agent_step(t, rs["meanWealth"])
        orc.compact(t)
        replaced = 0
        if c["agentReplacements"] > 0:
            orc.push_mt_to_python()
            replaced = pu.replace_dead_agents(c, state, pop, t)
            orc.set_mt_from_python()
        d, snap = pu.state_digests(state, *orc.mt_state())
        assert d == gold[t]["digests"], f"state differs at t={t}"
        rs = sb.update(orc.stats(t), t, replaced)
        for k in stats.LOG_KEYS:
            assert _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs[k]} vs {log[t][k]}"
