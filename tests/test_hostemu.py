"""The engine's mode E DEVICE rule source (csrc/ssb_rules.cuh, ssb_exact.cuh), compiled for the host as a
single lane (tests/hostemu), against the fixtures generated from the unmodified reference: every digest
of every step and all 59 log keys of the shipped examples the engine supports.  This puts the transaction
code nvcc compiles for sm_100a under the CPU suite; the `-m gpu` parity tests remain the proof for the
GPU build (lane-parallel ranking, device libm)."""
import pytest

import parity_util as pu
from hostemu_binding import HostEmu
from sugarscape_b200 import init, layout, stats
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


ETHICS_VARIANTS = ["ethics_trade", "ethics_combat", "ethics_replacement", "ethics_pollution",
                   # Bentham.updateSelfishnessFactor: the "Dynamic" names (default 0.01) and a configured range
                   "ethics_dynamic", "ethics_dynamic_configured",
                   # ethics.Temperance (simple variant), alone and mixed with Bentham agents and live combat
                   "ethics_temperance", "ethics_temperance_mixed",
                   # ... and with pecs=True (physical + emotional + cognitive + social scores)
                   "ethics_temperance_pecs", "ethics_temperance_pecs_combat",
                   # the "Top" orderings of the Bentham family
                   "ethics_top",
                   # ethics.Asimov among plain agents (ethics.py:8-98): with spice and births, sugar-only with live combat and
                   # replacement, and on top of every other family of determinism.json
                   "asimov_trade", "asimov_combat", "asimov_determinism", "asimov_temperance", "asimov_temperance_pecs", "asimov_bentham", "asimov_bentham_combat", "asimov_bentham_determinism",
                   # found by fuzzing: the initial population's stored neighbourhoods date from before the first infections
                   "asimov_bentham_radial_diseases"]


@pytest.mark.parametrize("variant", ETHICS_VARIANTS)
def test_device_ethics_source_matches_reference_trajectory(variant):
    """SURVEY row a21 on the engine's own source: the ethics.Bentham decision models (bentham, altruist,
    egoist, negativeBentham and the lookahead spawn variants) as the device code implements them
    (csrc/ssb_rules.cuh: ethical_pick / ethical_value, child inheritance; host: layout.HostEthics), against
    fixtures from the unmodified reference (tests/golden/make_golden_variants.py) on top of trade + tagging +
    reproduction, live combat, host-side replacement and pollution.  Every digest of every step and all 59
    log keys, the four decision-model keys from ethics_stats_body."""
    base, overrides, gold, log = pu.load_variant(variant)
    c, state, pop, params, endow, tags = pu.build(base, overrides=overrides)
    assert state.ethics is not None
    history = _run_with_extensions(c, state, pop, params, endow, tags, gold, log)
    # the ethical ranking overrode the greedy choice in (almost) every step
    assert sum(h["agentLastMoveOptimalityPercentage"] < 100 for h in history) > len(gold) * 0.9


def _run_with_extensions(c, state, pop, params, endow, tags, gold, log):
    """The stepping loop with every ABI 2 extension the configuration needs (decision models, lending, friends)."""
    emu = HostEmu(params, state, endow, tags)
    emu.set_mt_from_python()
    d, _ = pu.state_digests(state, *emu.mt_state())
    assert d == gold[0]["digests"], "t=0 state differs"
    sb = stats.StatsBuilder(c, init.equator(c))
    rs = sb.update(emu.stats(0), 0, 0, emu.ethics_stats(), emu.ext_stats(0))
    history = [dict(rs)]
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
        rs = sb.update(emu.stats(t), t, replaced, emu.ethics_stats(), emu.ext_stats(t))
        for k in stats.LOG_KEYS:
            assert _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs[k]} vs {log[t][k]}"
        history.append(dict(rs))
    return history


def test_device_lending_source_matches_reference_on_lending_basic():
    """examples/lending_basic.json through the device source of doLending / payDebt / updateLoans (csrc/ssb_ext.cuh:
    loan lists as chains through a pool with Python's list semantics): digests and all 59 log keys, loanVolume
    included; about 2.5 k loans in 200 steps."""
    c, state, pop, params, endow, tags = pu.build("lending_basic")
    assert state.ext is not None and "lending" in state.ext.families
    meta, _ = pu.load_golden("lending_basic")
    history = _run_with_extensions(c, state, pop, params, endow, tags, meta["steps"], meta["log"])
    assert sum(h["loanVolume"] for h in history) > 1000


def test_device_lending_friends_combat_source_matches_reference_on_dune():
    """examples/dune.json: lending with inheritance to children (payDebtToCreditorChildren), friends, live combat
    (conflict happiness), trade, tags, reproduction -- digests and all 59 log keys."""
    c, state, pop, params, endow, tags = pu.build("dune")
    assert state.ext is not None and state.ext.families == {"lending", "social"}
    meta, _ = pu.load_golden("dune")
    history = _run_with_extensions(c, state, pop, params, endow, tags, meta["steps"], meta["log"])
    assert sum(h["meanConflictHappiness"] != 0 for h in history) > 50


def test_device_disease_source_matches_reference_on_disease_basic():
    """examples/disease_basic.json through the device source of doDisease / catchDisease (csrc/ssb_ext.cuh) and the
    host's configureDiseases (init.configure_diseases): extinct after five steps, every digest and log key on the way."""
    c, state, pop, params, endow, tags = pu.build("disease_basic")
    assert state.ext is not None and "diseases" in state.ext.families
    meta, _ = pu.load_golden("disease_basic")
    history = _run_with_extensions(c, state, pop, params, endow, tags, meta["steps"], meta["log"])
    assert sum(h["sickAgents"] > 0 for h in history) >= 4


@pytest.mark.parametrize("variant,min_sick_steps", [("disease_mild", 15), ("disease_endemic", 80), ("disease_late_start", 25)])
def test_device_disease_source_matches_reference_on_variants(variant, min_sick_steps):
    """Milder diseases (variant fixtures from the unmodified reference): incubation, transmission / protection draws,
    recovery through the immune response, births with inherited immune systems, vision pushed to 0."""
    base, overrides, gold, log = pu.load_variant(variant)
    c, state, pop, params, endow, tags = pu.build(base, overrides=overrides)
    history = _run_with_extensions(c, state, pop, params, endow, tags, gold, log)
    assert sum(h["sickAgents"] > 0 for h in history) >= min_sick_steps


@pytest.mark.parametrize("variant", ["nowrap_trade_tagging_reproduction", "nowrap_pollution", "moore_trade_tagging_reproduction", "moore_pollution",
                                     "radial_trade_tagging_reproduction", "radial_combat", "radial_small_map", "radial_nowrap_small_map"])
def test_device_source_without_wraparound_matches_reference(variant):
    """environmentWraparound = false (cell.py:47-121; variant fixtures from the unmodified reference): tagging, trade and
    births over neighbour lists that stop at the map's edge, pollution flux over the neighbours that exist.  moore_*: eight
    neighbours (cell.py:77-89).  radial_*: discs instead of crosses (environment.py:146-173), with live combat and on a
    map smaller than the ranges."""
    base, overrides, gold, log = pu.load_variant(variant)
    c, state, pop, params, endow, tags = pu.build(base, overrides=overrides)
    assert params.flags & (layout.F_NOWRAP | layout.F_MOORE | layout.F_RADIAL)
    _run_with_extensions(c, state, pop, params, endow, tags, gold, log)


@pytest.mark.parametrize("variant,ethics", [("determinism_plain", False), ("determinism_ethics", True), ("nowrap_determinism", True), ("moore_nowrap_determinism", True),
                                            ("radial_determinism", True), ("radial_nowrap_determinism", True), ("determinism_late_diseases", True)])
def test_device_source_matches_reference_on_determinism_variants(variant, ethics):
    """examples/determinism.json (variant fixtures from the unmodified reference): diseases, depression, lending,
    friends, racial tags, universal income, live combat, trade, tags, inheritance, pollution and seasons at once --
    `_plain` without, `_ethics` with the ethics.Bentham decision models.  Only the experimental-group split of the
    log is switched off.  Every digest and all 59 log keys for 200 steps, on the device source."""
    base, overrides, gold, log = pu.load_variant(variant)
    c, state, pop, params, endow, tags = pu.build(base, overrides=overrides)
    assert state.ext.families == {"lending", "social", "diseases", "misc"} and (state.ethics is not None) == ethics
    history = _run_with_extensions(c, state, pop, params, endow, tags, gold, log)
    non_greedy = sum(h["agentLastMoveOptimalityPercentage"] < 100 for h in history)
    assert (non_greedy > 0.7 * (len(gold) - 1)) == ethics


@pytest.mark.parametrize("name", ["determinism", "all_features"])
def test_device_source_matches_reference_on_all_features_examples(name):
    """examples/determinism.json (the example BASELINE.json's north star names) and all_features.json AS SHIPPED on the
    device source: every rule family, the ethics decision models and the experimental group "depressed" -- every digest
    of every step and ALL 203 log keys (59 base keys, the neighbourhood split, the complete group / control copies
    with their interaction counters)."""
    c, state, pop, params, endow, tags = pu.build(name)
    meta, _ = pu.load_golden(name)
    gold, log = meta["steps"], meta["log"]
    assert state.ext.families == {"lending", "social", "diseases", "misc", "group"} and state.ethics is not None
    emu = HostEmu(params, state, endow, tags)
    emu.set_mt_from_python()
    d, _ = pu.state_digests(state, *emu.mt_state())
    assert d == gold[0]["digests"], "t=0 state differs"
    sb = stats.StatsBuilder(c, init.equator(c))
    assert len(sb.runtime_stats) == 203
    rs = sb.update_with_groups(emu.stats(0), 0, 0, emu.ethics_stats(), emu.ext_stats(0), emu.group_stats(0))
    for t in range(0, len(gold)):
        if t > 0:
            emu.env_step(t)
            emu.agent_step(t, rs["meanWealth"])
            emu.compact(t)
            d, snap = pu.state_digests(state, *emu.mt_state())
            assert d == gold[t]["digests"], f"state differs at t={t}"
            rs = sb.update_with_groups(emu.stats(t), t, 0, emu.ethics_stats(), emu.ext_stats(t), emu.group_stats(t))
        assert len(log[t]) == 203
        for k in log[t]:
            assert k in rs and _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs.get(k)} vs {log[t][k]}"


def test_dead_creditor_table_wraps_around():
    """The dead-creditor table of the lending family is a ring (include/ssb.h): with room for only 8 entries it wraps
    more than twice during dune.json's 200 steps (22 creditors die with loans outstanding) and the trajectory still
    equals the reference's, because an entry matters for at most the longest loan duration after the death."""
    c, state, pop, params, endow, tags = pu.build("dune")
    ext = state.ext
    ext.sizes["dead"] = 8
    for name in ("dead_slot", "dead_id", "dead_children_head", "dead_time"):
        ext.arrays["lending." + name] = ext.arrays["lending." + name][:8].copy()
    meta, _ = pu.load_golden("dune")
    _run_with_extensions(c, state, pop, params, endow, tags, meta["steps"], meta["log"])
    from sugarscape_b200 import layout
    assert int(ext.arrays["lending.counters"][layout.LC_N_DEAD]) > 2 * 8


@pytest.mark.parametrize("variant", ["determinism_group_female", "determinism_group_male", "determinism_group_race1", "determinism_group_bentham",
                                     "determinism_group_sick", "disease_mild_group_sick", "combat_group_male_replacement"])
def test_device_source_matches_reference_on_other_experimental_groups(variant):
    """determinism.json with another experimental group than the shipped "depressed" (variant fixtures of the unmodified
    reference): membership by sex, by race (race<N>) and by decision model (the group is a decisionModel string) -- digests
    and all 203 log keys.  (Groups whose membership changes during the run are refused.)"""
    base, overrides, gold, log = pu.load_variant(variant)
    c, state, pop, params, endow, tags = pu.build(base, overrides=overrides)
    assert "group" in state.ext.families and state.ext.group_kind > 1
    emu = HostEmu(params, state, endow, tags)
    emu.set_mt_from_python()
    sb = stats.StatsBuilder(c, init.equator(c))
    rs = sb.update_with_groups(emu.stats(0), 0, 0, emu.ethics_stats(), emu.ext_stats(0), emu.group_stats(0))
    for t in range(0, len(gold)):
        if t > 0:
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
            rs = sb.update_with_groups(emu.stats(t), t, replaced, emu.ethics_stats(), emu.ext_stats(t), emu.group_stats(t),
                                       state.newest_members(replaced))
        assert len(log[t]) == 203
        for k in log[t]:
            assert k in rs and _close(rs[k], log[t][k]), f"log key {k} differs at t={t}: {rs.get(k)} vs {log[t][k]}"


def _agentlog_cases():
    import gzip
    import json
    import os
    with gzip.open(os.path.join(pu.GOLDEN, "agentlog.json.gz"), "rt") as f:
        return json.load(f)


@pytest.mark.parametrize("case", ["determinism", "combat", "trade_pollution", "radial_asimov"])
def test_per_agent_log_matches_reference(case, tmp_path, monkeypatch):
    """`agentLogfile` (reference agent.py:1567-1620, sugarscape.py:415-421): the drop-in driver
    (sugarscape_b200.simulation.Sugarscape) on the host-compiled device source writes the per-agent log the unmodified
    reference writes -- same records in the same (action) order, every field equal (1e-9), incl. the quirks: the records
    of step 1 appear twice, agent.timeToLive is whatever the last findTimeToLive() call of anybody left."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.join(pu.ROOT, "tools"))
    from run_gpu_tests_on_hostemu import HostEmuEngine
    import sugarscape_b200.simulation as simulation
    monkeypatch.setattr(simulation, "Engine", HostEmuEngine)
    fixture = _agentlog_cases()[case]
    ov = dict(fixture["overrides"], agentLogfile=str(tmp_path / "agents.json"), logfile=str(tmp_path / "log.json"))
    c = pu.example_configuration(fixture["base"], ov)
    c["agentLogfile"], c["logfile"] = ov["agentLogfile"], ov["logfile"]
    simulation.Sugarscape(c).runSimulation(c["timesteps"])
    mine = json.load(open(ov["agentLogfile"]))
    ref = fixture["agent_log"]
    assert len(mine) == len(ref)
    for i, (a, b) in enumerate(zip(mine, ref)):
        assert list(a) == list(b), f"record {i}: keys differ"
        for k in b:
            assert _close(a[k], b[k]), f"record {i} (t={b['timestep']}, agent {b['ID']}): {k} = {a[k]} vs {b[k]}"
    log = json.load(open(ov["logfile"]))
    assert len(log) == len(fixture["log"])


def test_the_gpu_verified_examples_stay_on_the_plain_kernels():
    """The 24 examples that were verified on a B200 in round 1 bind no ABI 2 extension, so they run the EXT = false
    instantiations whose SASS is unchanged (profiles/r1_sass_equivalence_measured_vs_head.txt); the other five bind
    exactly the families they configure."""
    expected = {"lending_basic": {"lending"}, "dune": {"lending", "social"}, "disease_basic": {"diseases"},
                "determinism": {"lending", "social", "diseases", "misc", "group"},
                "all_features": {"lending", "social", "diseases", "misc", "group"}}
    for name in SUPPORTED + sorted(expected):
        c, state, pop, params, endow, tags = pu.build(name)
        families = state.ext.families if state.ext is not None else set()
        assert families == expected.get(name, set()), name
        assert (state.ethics is not None) == (name in ("determinism", "all_features")), name
