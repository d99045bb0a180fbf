#!/usr/bin/env python3
"""Fuzz the CPU oracle against the live, UNMODIFIED reference (build container only).

    python tests/golden/fuzz_oracle.py [cases] [seed]

Every case draws a base example, a map size, a population, a seed and a random set of rule
overrides inside the space the oracle restates, runs the reference and the oracle side by side and
compares the per-step digests (agents, fp64 holdings, tags, grid, pollution, MT19937 state).  The
first mismatch prints the overrides so that the case can be turned into a fixture
(make_golden_variants.py).  Development tool: the reference does not exist on the GPU box.
"""
import os
import random
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_harness as rh  # noqa: E402
import parity_util as pu  # noqa: E402
from oracle_binding import Oracle  # noqa: E402
from sugarscape_b200 import init, layout, stats  # noqa: E402

BASES = ["determinism", "dune", "trade_tagging_reproduction", "combat_unlimited", "lending_basic", "disease_basic",
         "pollution_formation", "reproduction_inheritance", "seasonal_migration", "cultural_combat_unlimited"]


def draw_overrides(rng, base):
    side = rng.choice([20, 24, 30, 36, 40, 50])
    ov = {"seed": rng.randrange(1, 10 ** 6), "environmentWidth": side, "environmentHeight": side,
          "startingAgents": rng.randrange(40, max(41, side * side // 5)), "timesteps": 60,
          "experimentalGroup": None}
    if rng.random() < 0.5:
        ov["agentVision"] = sorted([rng.randrange(1, 7), rng.randrange(1, 7)])
    if rng.random() < 0.5:
        ov["agentMovement"] = sorted([rng.randrange(1, 7), rng.randrange(1, 7)])
    if rng.random() < 0.4:
        ov["agentAggressionFactor"] = rng.choice([[0, 0], [0, 1], [1, 1], [0, 2]])
    if rng.random() < 0.4:
        ov["agentInheritancePolicy"] = rng.choice(["none", "children", "sons", "daughters"])
    if rng.random() < 0.3:
        ov["agentReplacements"] = rng.choice([0, ov["startingAgents"] // 2, ov["startingAgents"]])
    if rng.random() < 0.4:
        ov["agentLendingFactor"] = rng.choice([[0, 0], [1, 1], [0, 2]])
        ov["agentLoanDuration"] = sorted([rng.randrange(1, 8), rng.randrange(1, 8)])
        ov["agentBaseInterestRate"] = rng.choice([[0.05, 0.1], [0.1, 0.1], [0.5, 1.5]])
    if rng.random() < 0.4:
        ov["agentMaxFriends"] = rng.choice([[0, 0], [1, 3], [5, 10]])
    if rng.random() < 0.5:
        n = rng.choice([0, 3, 10, 30])
        ov.update({"startingDiseases": n, "startingDiseasesPerAgent": rng.choice([[0, 0], [0, 2], [1, 4]]),
                   "agentImmuneSystemLength": rng.choice([12, 24, 40]), "diseaseTagStringLength": rng.choice([[2, 6], [5, 11]]),
                   "diseaseSugarMetabolismPenalty": rng.choice([[0, 0], [0, 1], [0.5, 1.3]]),
                   "diseaseVisionPenalty": rng.choice([[0, 0], [-1, 1]]), "diseaseMovementPenalty": rng.choice([[0, 0], [-1, 0]]),
                   "diseaseFertilityPenalty": rng.choice([[0, 0], [-1, 0]]), "diseaseAggressionPenalty": rng.choice([[0, 0], [-1, 1]]),
                   "diseaseIncubationPeriod": rng.choice([[0, 0], [0, 3]]), "diseaseTransmissionChance": rng.choice([[1.0, 1.0], [0.2, 0.9]]),
                   "agentDiseaseProtectionChance": rng.choice([[0.0, 0.0], [0.0, 0.6]])})
        if rng.random() < 0.4 or os.environ.get("SSB_FUZZ_LATE_DISEASES") == "1":      # diseases that start during the run
            ov["diseaseTimeframe"] = rng.choice([[0, 20], [5, 30], [1, 1], [0, 100]])
            if ov["startingDiseases"] == 0:
                ov["startingDiseases"] = 10
    if rng.random() < 0.4:
        ov["agentDepressionPercentage"] = rng.choice([0, 0.05, 0.3])
    if rng.random() < 0.3:
        ov["agentUniversalSugar"] = rng.choice([[0, 0], [1, 1], [0, 2]])
        ov["agentUniversalSpice"] = rng.choice([[0, 0], [1, 1]])
        ov["environmentUniversalSugarIncomeInterval"] = rng.choice([1, 3])
        ov["environmentUniversalSpiceIncomeInterval"] = rng.choice([1, 2])
    if rng.random() < 0.4:
        ov["agentDecisionModels"] = rng.choice([["none"], ["bentham"], ["altruist", "egoist"], ["none", "negativeBentham"],
                                                ["bentham", "altruist", "egoist", "negativeBentham", "none"]])
        ov["agentDecisionModelFactor"] = rng.choice([[0, 0], [1, 1]])
        ov["agentDecisionModelLookaheadFactor"] = rng.choice([0, 0.5, 1])
    if rng.random() < 0.3:
        ov["agentRacialTagStringLength"] = rng.choice([0, 5, 11])
        ov["environmentMaxRaces"] = rng.choice([1, 2, 3])
    if rng.random() < 0.3:
        ov["environmentPollutionTimeframe"] = rng.choice([[0, 0], [0, -1], [10, 40]])
        ov["environmentPollutionDiffusionDelay"] = rng.choice([0, 1, 3])
    if rng.random() < 0.3:
        ov["environmentSeasonInterval"] = rng.choice([0, 5, 20])
        ov["environmentSeasonalGrowbackDelay"] = rng.choice([0, 2])
    # neighborhoodMode "moore" (cell.py:77-89) and environmentWraparound = false (cell.py:47-121); SSB_FUZZ_NEIGHBOURS=1: always one of them
    forced_nb = os.environ.get("SSB_FUZZ_NEIGHBOURS") == "1"
    pick = rng.random()
    if forced_nb:
        pick = pick * 0.3
    if pick < 0.1:
        ov["neighborhoodMode"] = "moore"
    elif pick < 0.2:
        ov["environmentWraparound"] = False
    elif pick < 0.3:
        ov["neighborhoodMode"] = "moore"
        ov["environmentWraparound"] = False
    # ethics.Asimov among plain agents (device source only: fuzz_hostemu.py; the oracle has no Asimov, fuzz_oracle.py skips it)
    if os.environ.get("SSB_FUZZ_ASIMOV") == "1" or rng.random() < 0.1:
        ov["agentDecisionModels"] = rng.choice([["asimov"], ["asimov", "none"], ["none", "none", "asimov"], ["asimov", "rawSugarscape", "asimov"],
                                                ["asimov", "bentham"], ["asimov", "egoist", "none", "altruist"], ["bentham", "asimov", "none"]])
        ov["agentDecisionModelFactor"] = rng.choice([[1, 1], [0, 1], [0.5, 1]])
    # radial vision / movement (environment.py:146-173); SSB_FUZZ_RADIAL=1 puts every case there, half of them on maps
    # smaller than the ranges (per-axis caps, odd and even axes)
    forced = os.environ.get("SSB_FUZZ_RADIAL") == "1"
    if forced or rng.random() < 0.15:
        ov["agentVisionMode"] = ov["agentMovementMode"] = "radial"
        if forced and rng.random() < 0.5:
            ov["environmentWidth"], ov["environmentHeight"] = rng.choice([7, 8, 9, 10, 12, 13]), rng.choice([7, 8, 9, 11, 12, 20])
            ov["startingAgents"] = rng.randrange(5, ov["environmentWidth"] * ov["environmentHeight"] // 3)
            if "agentReplacements" in ov:
                ov["agentReplacements"] = min(ov["agentReplacements"], ov["startingAgents"])
    return ov


def run_case(base, ov, steps):
    path = os.path.join(rh.REFERENCE_DIR, "examples", base + ".json")
    try:
        S = rh.build_reference(path, ov)
        S.updateRuntimeStats()
        ref_digests = [rh.digests(rh.snapshot(S))]
        for _ in range(steps):
            if len(S.agents) == 0 and not S.keepAlive:
                break
            S.doTimestep()
            ref_digests.append(rh.digests(rh.snapshot(S)))
    except Exception as e:
        return "skipped (the reference itself raised %r)" % (e,)
    c, state, pop, params, endow, tags = pu.build(base, overrides=ov, check=False)
    if c["experimentalGroup"] is not None or c["agentLeader"]:
        return "skipped"
    if any("asimov" in m or "temperance" in m for m in c["agentDecisionModels"]):
        return "skipped (the oracle restates the Bentham family only; Asimov and Temperance are pinned on the device source, fuzz_hostemu.py)"
    if c["agentReplacements"] > 0 and ((c["agentTagging"] and c["agentTagStringLength"] > 0 and not c["environmentTribePerQuadrant"])
                                       or (c["agentImmuneSystemLength"] > 0 and c["startingDiseases"] > 0)):
        # Known deviation: the reference shares the tag / immune-system LIST objects between the
        # agents created from one endowment entry
        return "skipped (reference aliases list-valued endowments between replacement twins)"
    orc = Oracle(params, state, endow, tags)
    horizon = steps + 8
    orc.bind_lending(pop.endowments)
    orc.bind_social(pop.endowments)
    orc.bind_misc(c, pop.endowments, horizon)
    orc.bind_ethics(c, pop.endowments)
    if c["agentImmuneSystemLength"] > 0:
        orc.bind_disease(c, pop.endowments, pop.disease_endowments, horizon)
    orc.set_mt_from_python()
    try:
        d, _ = pu.state_digests(state, *orc.mt_state())
        if d != ref_digests[0]:
            return "t=0 differs: " + str({k: d[k] == ref_digests[0][k] for k in d})
        sb = stats.StatsBuilder(c, init.equator(c))
        rs = sb.update(orc.stats(0), 0)
        for t in range(1, len(ref_digests)):
            orc.env_step(t)
            orc.agent_step(t, rs["meanWealth"])
            orc.compact(t)
            if c["agentReplacements"] > 0:          # replaceDeadAgents on the host, with the main stream (exact mode)
                n = int(state.counters[layout.C_N_LIST])
                if n < c["agentReplacements"]:
                    orc.push_mt_to_python()
                    new_agents = pop.place(c["agentReplacements"] - n, state.occ >= 0, t, int(state.counters[layout.C_NEXT_ID]))
                    state.append_agents(new_agents, protect_last=int(state.counters[layout.C_N_DEAD]))
                    state.counters[layout.C_NEXT_ID] += len(new_agents)
                    orc.adopt_new_agents(new_agents, c)
                    orc.set_mt_from_python()
            d, _ = pu.state_digests(state, *orc.mt_state())
            if d != ref_digests[t]:
                return f"t={t} differs: " + str({k: d[k] == ref_digests[t][k] for k in d})
            rs = sb.update(orc.stats(t), t, 0)
        return "ok %d steps, final population %d" % (len(ref_digests) - 1, rs["population"])
    finally:
        orc.unbind_lending(); orc.unbind_social(); orc.unbind_misc(); orc.unbind_ethics()
        if c["agentImmuneSystemLength"] > 0:
            orc.unbind_disease()


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    bad = 0
    for i in range(cases):
        base = rng.choice(BASES)
        ov = draw_overrides(rng, base)
        try:
            res = run_case(base, ov, 60)
        except Exception as e:
            res = "EXCEPTION " + repr(e) + "\n" + traceback.format_exc(limit=4)
        flag = "" if res.startswith("ok") or res.startswith("skipped") else "   <<<<<<"
        print(f"case {i} {base}: {res}{flag}", flush=True)
        if flag:
            bad += 1
            print("   overrides:", ov, flush=True)
    print("mismatching cases:", bad)


if __name__ == "__main__":
    main()
