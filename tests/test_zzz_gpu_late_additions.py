"""SURVEY row (f)3 on the GPU, the two pieces written after round 2's GPU minutes were spent.

(1) agentVisionMode = agentMovementMode = "radial" (reference environment.py:146-173, agent.py:766-779)
through the C ABI against six fixtures of the unmodified reference (tests/golden/make_golden_variants.py, radial_*; two of them without wraparound): the candidate
cells of a move are the disc floor(distance) <= range in the insertion order of Environment.findRadialCellRanges, for plain
agents (tagging + trade + births, live combat with retaliators and host-side replacement, a map smaller than the ranges) and
for every family of determinism.json incl. the Bentham decision models.  Integer state, tags, grid and the MT19937 stream
bit-exact on every step, fp64 holdings through their digest, all 59 log keys.

(2) ethics.Asimov among plain agents, Temperance agents (both variants) and the Bentham family (reference ethics.py:8-98; `asimov_pick` in csrc/ssb_rules.cuh)
against nine fixtures of the unmodified reference (asimov_*): with spice and births, sugar-only with live combat and replacement,
on top of every other family of determinism.json, and next to Temperance and Bentham agents that answer the second law (the
latter over the neighbourhood lists they stored at their own last ranking, at birth or at the last replacement).

(3) experimental groups other than the shipped "depressed": membership by sex, by race (race<N>), by decision model (the
group is a decisionModel string) and -- changing during the run -- the sick, on determinism.json and on a disease example --
digests and all 203 log keys against six fixtures of the unmodified reference (*_group_*).

(4) diseases that start during the run (diseaseTimeframe; Sugarscape.addRemainingDiseases, sugarscape.py:139-151): two fixtures
(*late*), oracle and device source.

Status: the six radial fixtures pass on the CPU oracle (tests/test_oracle_golden.py) and on the device rule source compiled for
the host (tests/test_hostemu.py), the nine Asimov fixtures on the latter (the oracle restates the Bentham family only); the
four group fixtures on the latter as well (the oracle's group is "depressed"); these cases have NOT run on a B200 yet.  The file sorts after every other test file so that its
first GPU run cannot hide anything else."""
import pytest

import parity_util as pu
from sugarscape_b200 import init, layout, stats
from test_gpu_parity import _close, _engine, _replace
from test_zz_gpu_decision_models import _Budget, _cap
from test_zz_gpu_decision_models import test_drop_in_driver_writes_the_reference_per_agent_log as _per_agent_log_case

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("variant", ["radial_small_map", "radial_trade_tagging_reproduction", "radial_combat", "radial_determinism",
                                     "radial_nowrap_small_map", "radial_nowrap_determinism",
                                     "asimov_trade", "asimov_combat", "asimov_determinism", "asimov_temperance", "asimov_temperance_pecs",
                                     "asimov_bentham", "asimov_bentham_combat", "asimov_bentham_determinism", "asimov_bentham_radial_diseases",
                                     "disease_late_start", "determinism_late_diseases"])
def test_radial_ranges_and_asimov_match_reference_trajectory(variant):
    base, overrides, gold, log = pu.load_variant(variant)
    c, state, pop, params, endow, tags, eng = _engine(base, layout.RNG_EXACT, overrides)
    assert ((params.flags & layout.F_RADIAL) if variant.startswith("radial") else (state.ethics.arrays["model"] == 4).any() if variant.startswith("asimov")
            else c["diseaseTimeframe"] != [0, 0])
    sb = stats.StatsBuilder(c, init.equator(c))
    eng.reduce_stats(0)
    rs = sb.update(eng.stats(), 0, 0, eng.ethics_stats(), eng.ext_stats())
    gold, budget = _cap(gold), _Budget()
    for t in range(1, len(gold)):
        if budget.spent(t - 1):
            break
        eng.env_step(t)
        eng.agent_step(t, rs["meanWealth"])
        eng.compact(t)
        replaced = _replace(c, eng, pop, t) if c["agentReplacements"] > 0 else 0
        eng.reduce_stats(t)
        rs = sb.update(eng.stats(), t, replaced, eng.ethics_stats(), eng.ext_stats())
        d, snap = pu.state_digests(eng.download(), *eng.mt_state())
        assert d == gold[t]["digests"], f"{variant}: state differs from the reference at t={t}"
        for k in stats.LOG_KEYS:
            assert _close(rs[k], log[t][k]), f"{variant}: log key {k} at t={t}: {rs[k]} vs {log[t][k]}"
    eng.close()


@pytest.mark.parametrize("variant", ["determinism_group_female", "determinism_group_male", "determinism_group_race1", "determinism_group_bentham",
                                     "determinism_group_sick", "disease_mild_group_sick", "combat_group_male_replacement"])
def test_other_experimental_groups_match_reference_trajectory_and_full_log(variant):
    base, overrides, gold, log = pu.load_variant(variant)
    c, state, pop, params, endow, tags, eng = _engine(base, layout.RNG_EXACT, overrides)
    assert "group" in state.ext.families and state.ext.group_kind > 1
    sb = stats.StatsBuilder(c, init.equator(c))
    eng.reduce_stats(0)
    rs = sb.update_with_groups(eng.stats(), 0, 0, eng.ethics_stats(), eng.ext_stats(), eng.group_stats())
    gold, budget = _cap(gold), _Budget()
    for t in range(0, len(gold)):
        if budget.spent(t - 1):
            break
        if t > 0:
            eng.env_step(t)
            eng.agent_step(t, rs["meanWealth"])
            eng.compact(t)
            replaced = members = 0
            if c["agentReplacements"] > 0 and int(eng.counters()[layout.C_N_LIST]) < c["agentReplacements"]:
                hs = eng.download()              # (test_gpu_parity._replace, keeping the host state for the membership count)
                eng.push_mt_to_python()
                replaced = pu.replace_dead_agents(c, hs, pop, t)
                eng.upload(hs, [key for key in eng.tensors if key != "pollution_flux"])
                eng.set_mt_from_python()
                members = hs.newest_members(replaced)
            eng.reduce_stats(t)
            rs = sb.update_with_groups(eng.stats(), t, replaced, eng.ethics_stats(), eng.ext_stats(), eng.group_stats(), members)
            d, snap = pu.state_digests(eng.download(), *eng.mt_state())
            assert d == gold[t]["digests"], f"{variant}: state differs from the reference at t={t}"
        assert len(log[t]) == 203
        for k in log[t]:
            assert k in rs and _close(rs[k], log[t][k]), f"{variant}: log key {k} at t={t}: {rs.get(k)} vs {log[t][k]}"
    eng.close()


def test_drop_in_driver_writes_the_per_agent_log_with_radial_ranges_and_asimov(tmp_path):
    """`agentLogfile` through `Sugarscape(conf).runSimulation()` on determinism.json with radial ranges and Asimov robots among
    plain agents (tests/golden/make_golden_agentlog.py, case radial_asimov): record by record."""
    _per_agent_log_case("radial_asimov", tmp_path)
