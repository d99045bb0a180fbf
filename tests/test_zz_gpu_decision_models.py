"""SURVEY row a21 and the first (f)2 families on the GPU: the ethics.Bentham decision models, lending and
friends through the C ABI (ssb_bind_ethics / ssb_bind_ext, the extended instantiation of the mode E kernel) against fixtures generated from the unmodified reference
(tests/golden/make_golden_variants.py: ethics on top of trade + tagging + reproduction, live combat, host-side
replacement, pollution).  Integer state, tags, grid and the MT19937 stream bit-exact on every step, fp64 holdings
through their digest, all 59 log keys.

Every case replays its WHOLE fixture (200 timesteps) and fails if it did not: the first B200 run of these kernels
(round 1, as xfail(strict=False) cases) passed, so the marker is gone.  The file sorts last so that nothing runs after
it in the same process."""
import os
import time

import pytest

import parity_util as pu
from sugarscape_b200 import init, layout, stats
from test_gpu_parity import _close, _engine, _replace

pytestmark = [pytest.mark.gpu]

# The one-warp extended kernel walks Python-list semantics on a single lane; a 200-step replay with a state download and
# digests after every step takes some tens of seconds.  SSB_GPU_EXT_STEPS > 0 caps the replay for quick local runs (the
# default 0 replays everything); SSB_GPU_EXT_SECONDS is a safety net per case -- a case that hits it FAILS (it did not
# cover its fixture), it does not pass on a prefix.
STEPS = int(os.environ.get("SSB_GPU_EXT_STEPS", "0"))
SECONDS = float(os.environ.get("SSB_GPU_EXT_SECONDS", "240"))
DRIVER_STEPS = 30           # (with SSB_GPU_EXT_STEPS > 0 only) the drop-in driver runs cannot be interrupted: a fixed, short run


def _cap(gold):
    return gold if STEPS <= 0 else gold[:STEPS + 1]


class _Budget:
    def __init__(self):
        self.t0 = time.time()
        self.ran = 0

    def spent(self, t):
        self.ran = t
        if SECONDS > 0 and t >= 10 and time.time() - self.t0 > SECONDS:
            pytest.fail(f"only {t} timesteps compared in {SECONDS:.0f} s: the case did not cover its fixture")
        return False


ETHICS_VARIANTS = ["ethics_trade", "ethics_combat", "ethics_replacement", "ethics_pollution", "ethics_dynamic",
                   "ethics_dynamic_configured", "ethics_temperance", "ethics_temperance_mixed",
                   "ethics_temperance_pecs", "ethics_temperance_pecs_combat", "ethics_top"]


@pytest.mark.parametrize("variant", ETHICS_VARIANTS)
def test_decision_models_match_reference_trajectory(variant):
    base, overrides, gold, log = pu.load_variant(variant)
    c, state, pop, params, endow, tags, eng = _engine(base, layout.RNG_EXACT, overrides)
    assert state.ethics is not None
    sb = stats.StatsBuilder(c, init.equator(c))
    eng.reduce_stats(0)
    rs = sb.update(eng.stats(), 0, 0, eng.ethics_stats(), eng.ext_stats())
    non_greedy = 0
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
        non_greedy += rs["agentLastMoveOptimalityPercentage"] < 100
        for k in stats.LOG_KEYS:
            assert _close(rs[k], log[t][k]), f"{variant}: log key {k} at t={t}: {rs[k]} vs {log[t][k]}"
    assert non_greedy > 0.8 * budget.ran
    eng.close()


@pytest.mark.parametrize("name", ["lending_basic", "dune"])
def test_lending_and_friends_examples_match_reference_trajectory(name):
    """examples/lending_basic.json and examples/dune.json (lending incl. payDebtToCreditorChildren, friends, conflict
    happiness on top of trade, tags, live combat, reproduction and inheritance) through ssb_bind_ext: every digest
    of the 200-step protocol and all 59 log keys against the fixtures of the unmodified reference."""
    c, state, pop, params, endow, tags, eng = _engine(name)
    assert state.ext is not None
    meta, _ = pu.load_golden(name)
    gold, log = meta["steps"], meta["log"]
    sb = stats.StatsBuilder(c, init.equator(c))
    eng.reduce_stats(0)
    rs = sb.update(eng.stats(), 0, 0, eng.ethics_stats(), eng.ext_stats())
    volume = 0
    gold, budget = _cap(gold), _Budget()
    for t in range(1, len(gold)):
        if budget.spent(t - 1):
            break
        eng.env_step(t)
        eng.agent_step(t, rs["meanWealth"])
        eng.compact(t)
        eng.reduce_stats(t)
        rs = sb.update(eng.stats(), t, 0, eng.ethics_stats(), eng.ext_stats())
        d, snap = pu.state_digests(eng.download(), *eng.mt_state())
        assert d == gold[t]["digests"], f"{name}: state differs from the reference at t={t}"
        volume += rs["loanVolume"]
        for k in stats.LOG_KEYS:
            assert _close(rs[k], log[t][k]), f"{name}: log key {k} at t={t}: {rs[k]} vs {log[t][k]}"
    assert volume > 0 or budget.ran < 60          # (loans start once agents reach fertility)
    eng.close()


@pytest.mark.parametrize("case", ["disease_basic", "disease_mild", "disease_endemic"])
def test_disease_cases_match_reference_trajectory(case):
    """examples/disease_basic.json and two milder variants (fixtures of the unmodified reference): configureDiseases
    on the host, doDisease / catchDisease, the find* modifiers, inherited immune systems on the device."""
    if case == "disease_basic":
        meta, _ = pu.load_golden(case)
        base, overrides, gold, log = case, None, meta["steps"], meta["log"]
    else:
        base, overrides, gold, log = pu.load_variant(case)
    c, state, pop, params, endow, tags, eng = _engine(base, layout.RNG_EXACT, overrides)
    assert state.ext is not None and "diseases" in state.ext.families
    sb = stats.StatsBuilder(c, init.equator(c))
    eng.reduce_stats(0)
    rs = sb.update(eng.stats(), 0, 0, eng.ethics_stats(), eng.ext_stats())
    for k in ("sickAgents", "diseaseIncidence", "diseasePrevalence"):
        assert rs[k] == log[0][k], f"{k} differs after init"
    sick_steps = 0
    gold, budget = _cap(gold), _Budget()
    for t in range(1, len(gold)):
        if budget.spent(t - 1):
            break
        eng.env_step(t)
        eng.agent_step(t, rs["meanWealth"])
        eng.compact(t)
        eng.reduce_stats(t)
        rs = sb.update(eng.stats(), t, 0, eng.ethics_stats(), eng.ext_stats())
        d, snap = pu.state_digests(eng.download(), *eng.mt_state())
        assert d == gold[t]["digests"], f"{case}: state differs from the reference at t={t}"
        sick_steps += rs["sickAgents"] > 0
        for k in stats.LOG_KEYS:
            assert _close(rs[k], log[t][k]), f"{case}: log key {k} at t={t}: {rs[k]} vs {log[t][k]}"
    assert sick_steps >= min(4, budget.ran)
    eng.close()


@pytest.mark.parametrize("variant,ethics", [("determinism_plain", False), ("determinism_ethics", True), ("nowrap_determinism", True),
                                            ("nowrap_trade_tagging_reproduction", False), ("nowrap_pollution", False),
                                            ("moore_nowrap_determinism", True), ("moore_trade_tagging_reproduction", False), ("moore_pollution", False)])
def test_determinism_variants_match_reference_trajectory(variant, ethics):
    """examples/determinism.json without / with the ethics models, experimental-group split off (variant fixtures of the
    unmodified reference): every rule family at once.  nowrap_*: environmentWraparound = false (cell.py:47-121) on
    determinism.json, on tagging + trade + births and on pollution diffusion."""
    base, overrides, gold, log = pu.load_variant(variant)
    c, state, pop, params, endow, tags, eng = _engine(base, layout.RNG_EXACT, overrides)
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
        eng.reduce_stats(t)
        rs = sb.update(eng.stats(), t, 0, eng.ethics_stats(), eng.ext_stats())
        d, snap = pu.state_digests(eng.download(), *eng.mt_state())
        assert d == gold[t]["digests"], f"{variant}: state differs from the reference at t={t}"
        for k in stats.LOG_KEYS:
            assert _close(rs[k], log[t][k]), f"{variant}: log key {k} at t={t}: {rs[k]} vs {log[t][k]}"
    eng.close()


@pytest.mark.parametrize("name", ["determinism", "all_features"])
def test_all_features_examples_match_reference_trajectory_and_full_log(name):
    """examples/determinism.json (the example BASELINE.json names) and all_features.json as shipped: every digest of the
    200-step protocol and ALL 203 log keys, experimental group "depressed" included."""
    c, state, pop, params, endow, tags, eng = _engine(name)
    meta, _ = pu.load_golden(name)
    gold, log = meta["steps"], meta["log"]
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
            eng.reduce_stats(t)
            rs = sb.update_with_groups(eng.stats(), t, 0, eng.ethics_stats(), eng.ext_stats(), eng.group_stats())
            d, snap = pu.state_digests(eng.download(), *eng.mt_state())
            assert d == gold[t]["digests"], f"{name}: state differs from the reference at t={t}"
        assert len(log[t]) == 203
        for k in log[t]:
            assert k in rs and _close(rs[k], log[t][k]), f"{name}: log key {k} at t={t}: {rs.get(k)} vs {log[t][k]}"
    eng.close()


def test_drop_in_driver_writes_the_reference_log_for_determinism(tmp_path):
    """`Sugarscape(conf).runSimulation()` on examples/determinism.json as shipped: the 203-key log.json."""
    import json
    from sugarscape_b200.simulation import Sugarscape
    c = pu.example_configuration("determinism")
    c["logfile"] = str(tmp_path / "log.json")
    if STEPS > 0:
        c["timesteps"] = min(c["timesteps"], DRIVER_STEPS)
    Sugarscape(c).runSimulation(c["timesteps"])
    mine = json.load(open(c["logfile"]))
    meta, _ = pu.load_golden("determinism")
    log = meta["log"][:c["timesteps"] + 1]
    assert len(mine) == len(log) and len(mine[1]) == 203
    for entry, ref in zip(mine[1:], log[1:]):
        for k in ref:
            if entry is mine[-1] and k == "environmentWealthCreated":
                continue
            assert _close(entry[k], ref[k]), f"log key {k} at t={entry['timestep']}: {entry[k]} vs {ref[k]}"


@pytest.mark.parametrize("case", ["determinism", "combat", "trade_pollution"])
def test_drop_in_driver_writes_the_reference_per_agent_log(case, tmp_path):
    """`agentLogfile`: the per-agent log of the drop-in driver equals the unmodified reference's, record by record
    (tests/golden/make_golden_agentlog.py)."""
    import gzip
    import json
    import os
    from sugarscape_b200.simulation import Sugarscape
    with gzip.open(os.path.join(pu.GOLDEN, "agentlog.json.gz"), "rt") as f:
        fixture = json.load(f)[case]
    ov = dict(fixture["overrides"], agentLogfile=str(tmp_path / "agents.json"), logfile=str(tmp_path / "log.json"))
    c = pu.example_configuration(fixture["base"], ov)
    c["agentLogfile"], c["logfile"] = ov["agentLogfile"], ov["logfile"]
    Sugarscape(c).runSimulation(c["timesteps"])
    mine, ref = json.load(open(ov["agentLogfile"])), fixture["agent_log"]
    assert len(mine) == len(ref)
    for i, (a, b) in enumerate(zip(mine, ref)):
        for k in b:
            assert _close(a[k], b[k]), f"record {i} (t={b['timestep']}, agent {b['ID']}): {k} = {a[k]} vs {b[k]}"


def test_drop_in_driver_writes_the_reference_log_with_decision_models(tmp_path):
    """The reference-facing entry point (sugarscape_b200.simulation.Sugarscape, what `sugarscape.py --conf` runs)
    on the ethics_trade configuration: the log.json it writes equals the reference's log entry by entry."""
    import json
    from sugarscape_b200.simulation import Sugarscape
    base, overrides, gold, log = pu.load_variant("ethics_trade")
    c = pu.example_configuration(base, overrides)
    c["logfile"] = str(tmp_path / "log.json")
    if STEPS > 0:
        c["timesteps"] = min(c["timesteps"], DRIVER_STEPS)
        log = log[:c["timesteps"] + 1]
    S = Sugarscape(c)
    S.runSimulation(c["timesteps"])
    mine = json.load(open(c["logfile"]))
    # the file = the all-zero entry, then t = 1 .. last (the fixture's "log" holds the post-update stats from t = 0 on)
    assert mine[0]["population"] == 0 and mine[0]["timestep"] == 0
    assert len(mine) == len(log)
    for entry, ref in zip(mine[1:], log[1:]):
        for k in stats.LOG_KEYS:
            if entry is mine[-1] and k == "environmentWealthCreated":
                continue          # endLog recomputes it (covered by the shipped-example log tests)
            assert _close(entry[k], ref[k]), f"log key {k} at t={entry['timestep']}: {entry[k]} vs {ref[k]}"
