#!/usr/bin/env python3
"""Extra fixtures from the UNMODIFIED reference: shipped examples with option overrides, for rule
families whose shipped example exercises them too briefly (disease_basic is extinct after 5 steps).

    python tests/golden/make_golden_variants.py [variant ...]        (build container only)

Writes <variant>.json.gz = {"meta", "base", "overrides", "steps" (per-step digests), "log"}.
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
import ref_harness as rh  # noqa: E402

_ETHICS = {"agentDecisionModels": ["none", "bentham", "altruist", "egoist", "negativeBentham"],
           "agentDecisionModelFactor": [1, 1], "agentDecisionModelLookaheadDiscount": [0.5, 0.5],
           "agentDecisionModelLookaheadFactor": 0.5, "agentSelfishnessFactor": [-1, -1]}

_MILD = {"diseaseSugarMetabolismPenalty": [0, 1], "diseaseSpiceMetabolismPenalty": [0, 0],
         "diseaseVisionPenalty": [-1, 1], "diseaseFertilityPenalty": [-1, 0], "diseaseIncubationPeriod": [0, 3],
         "diseaseTransmissionChance": [0.3, 1.0], "agentDiseaseProtectionChance": [0.0, 0.5],
         "startingDiseases": 12, "startingDiseasesPerAgent": [0, 2], "diseaseTagStringLength": [3, 8],
         "agentImmuneSystemLength": 20}

VARIANTS = {
    # diseases that incubate, spread with a chance, can be resisted and are recovered from
    "disease_mild": ("disease_basic", 150, dict(_MILD)),
    # long tag strings against a short immune system: slow recovery, the diseases stay around
    # diseases that start during the run (diseaseTimeframe): addRemainingDiseases hands each one to the first agent of the
    # shuffled list that is not immune, from its start timestep on
    "disease_late_start": ("disease_basic", 150, dict(_MILD, diseaseTimeframe=[0, 40], startingDiseases=20)),
    "determinism_late_diseases": ("determinism", 100, {"diseaseTimeframe": [0, 30], "experimentalGroup": None}),
    "disease_endemic": ("disease_basic", 120, {
        "diseaseSugarMetabolismPenalty": [0, 1], "diseaseSpiceMetabolismPenalty": [0, 0], "diseaseVisionPenalty": [-1, 0],
        "diseaseIncubationPeriod": [0, 2], "diseaseTransmissionChance": [0.6, 1.0], "agentDiseaseProtectionChance": [0.0, 0.3],
        "startingDiseases": 20, "startingDiseasesPerAgent": [1, 4], "diseaseTagStringLength": [10, 16],
        "agentImmuneSystemLength": 24}),
    # determinism.json without the ethics decision models and the experimental-group log: diseases,
    # depression, lending, friends, racial tags, universal income, combat, trade, tags, inheritance,
    # pollution and seasons all at once
    "determinism_plain": ("determinism", 200, {"agentDecisionModels": ["none"], "agentDecisionModelFactor": [0, 0],
                                               "experimentalGroup": None}),
    # determinism.json with the ethics decision models (bentham, altruist, egoist, negativeBentham), only the
    # experimental-group split of the log switched off
    "determinism_ethics": ("determinism", 200, {"experimentalGroup": None}),
    # the map comes from a file (tests/golden/environment_30.json) instead of the peak formula; note that the
    # reference swaps the file's sugar and spice
    "environment_file": ("trade_tagging_reproduction", 120, {"environmentFile": os.path.join(HERE, "environment_30.json"),
                                                             "environmentWidth": 30, "environmentHeight": 30,
                                                             "startingAgents": 150, "environmentMaxSugar": 4, "environmentMaxSpice": 4}),
    # ---- ethics.Bentham decision models on top of rule families the ENGINE supports ----
    # trade + tagging + reproduction: children inherit the decision model of one parent
    "ethics_trade": ("trade_tagging_reproduction", 150, dict(_ETHICS)),
    # live combat (prey, retaliators and loot enter the welfare the ethical ranking re-scores)
    "ethics_combat": ("combat_limited", 150, dict(_ETHICS)),
    # host-side replacement of the dead + the lookahead spawn variants + a selfishness range for bentham
    "ethics_replacement": ("agent_replacement", 120, dict(_ETHICS, agentDecisionModels=[
        "none", "benthamHalfLookahead", "altruistNoLookahead", "egoist", "negativeBenthamNoLookahead", "bentham"],
        agentSelfishnessFactor=[0.2, 0.8])),
    # pollution enters the intensity term of the ethical value
    "ethics_pollution": ("trade_pollution", 120, dict(_ETHICS)),
    # dynamic selfishness (Bentham.updateSelfishnessFactor): the "Dynamic" names with the default 0.01, and a configured range
    "ethics_dynamic": ("trade_tagging_reproduction", 120, dict(_ETHICS, agentDecisionModels=[
        "none", "benthamDynamic", "egoistDynamic", "altruistNoLookaheadDynamic", "negativeBentham"])),
    "ethics_dynamic_configured": ("combat_limited", 100, dict(_ETHICS, agentDynamicSelfishnessFactor=[0.05, 0.1],
                                                              agentSelfishnessFactor=[0.3, 0.7], agentDecisionModels=["bentham", "none"])),
    # ethics.Temperance, simple variant: virtue rolls on the main stream, the temperance factor moves with them
    "ethics_temperance": ("trade_tagging_reproduction", 120, dict(_ETHICS, agentDecisionModels=["none", "temperance"],
                                                                  agentDecisionModelFactor=[0.3, 0.9],
                                                                  agentDynamicDecisionModelFactor=[0.05, 0.1])),
    "ethics_temperance_mixed": ("combat_limited", 100, dict(_ETHICS, agentDecisionModels=["bentham", "temperance", "none", "egoistDynamic"],
                                                            agentDecisionModelFactor=[0.5, 1], agentDynamicDecisionModelFactor=[0.0, 0.2],
                                                            agentAggressionFactor=[0, 2], agentVision=[1, 3])),
    # Temperance with the physical / emotional / cognitive / social scores (pecs=True) next to the simple variant
    "ethics_temperance_pecs": ("trade_tagging_reproduction", 120, dict(_ETHICS, agentDecisionModels=["none", "temperancePECS", "temperance"],
                                                                       agentDynamicSocialPressureFactor=[0.01, 0.1],
                                                                       agentDynamicDecisionModelFactor=[0.0, 0.1])),
    "ethics_temperance_pecs_combat": ("combat_limited", 100, dict(_ETHICS, agentDecisionModels=["temperancePECS", "bentham"],
                                                                  agentDynamicSocialPressureFactor=[0.05, 0.05],
                                                                  agentAggressionFactor=[0, 2], agentVision=[2, 5])),
    # ---- environmentWraparound = false (cell.py:47-121): no neighbours across the map's edge, the reference's own test for
    # the south neighbour included; tagging + trade + births, pollution diffusion, and every family of determinism.json
    "nowrap_trade_tagging_reproduction": ("trade_tagging_reproduction", 150, {"environmentWraparound": False}),
    "nowrap_pollution": ("trade_pollution", 120, {"environmentWraparound": False}),
    "nowrap_determinism": ("determinism", 120, {"environmentWraparound": False, "experimentalGroup": None}),
    # ---- neighborhoodMode = "moore" (cell.py:77-89): eight neighbour cells in the order N, S, E, W, NE, NW, SE, SW -- alone
    # and together with environmentWraparound = false
    "moore_trade_tagging_reproduction": ("trade_tagging_reproduction", 150, {"neighborhoodMode": "moore"}),
    "moore_pollution": ("trade_pollution", 120, {"neighborhoodMode": "moore"}),
    "moore_nowrap_determinism": ("determinism", 120, {"neighborhoodMode": "moore", "environmentWraparound": False, "experimentalGroup": None}),
    # ---- agentVisionMode = agentMovementMode = "radial" (environment.py:146-173): discs of floor(distance) <= range, the
    # rings in the insertion order of findRadialCellRanges -- tagging + trade + births, live combat with retaliators, and
    # every family of determinism.json incl. the Bentham models (disease bonuses to vision / movement switched off: they
    # would push agents beyond the engine's candidate arrays)
    "radial_trade_tagging_reproduction": ("trade_tagging_reproduction", 120, {"agentVisionMode": "radial", "agentMovementMode": "radial"}),
    "radial_combat": ("combat_limited", 120, {"agentVisionMode": "radial", "agentMovementMode": "radial"}),
    "radial_determinism": ("determinism", 100, {"agentVisionMode": "radial", "agentMovementMode": "radial", "experimentalGroup": None,
                                                "diseaseVisionPenalty": [-1, 0], "diseaseMovementPenalty": [-1, 0]}),
    # radial ranges without wraparound: plain distances, discs clipped at the map's edge (environment.py:141-144, 175-179) -- on
    # every family of determinism.json (with the Moore neighbourhood on top) and on a map smaller than the ranges
    "radial_nowrap_determinism": ("determinism", 100, {"agentVisionMode": "radial", "agentMovementMode": "radial", "experimentalGroup": None,
                                                       "environmentWraparound": False, "neighborhoodMode": "moore",
                                                       "diseaseVisionPenalty": [-1, 0], "diseaseMovementPenalty": [-1, 0]}),
    "radial_nowrap_small_map": ("trade_tagging_reproduction", 100, {"agentVisionMode": "radial", "agentMovementMode": "radial",
                                                                    "environmentWraparound": False, "environmentWidth": 9, "environmentHeight": 12,
                                                                    "startingAgents": 30, "agentVision": [1, 6]}),
    # a map smaller than the ranges: the per-axis caps (width // 2) and the coinciding ends of an even axis
    "radial_small_map": ("trade_tagging_reproduction", 100, {"agentVisionMode": "radial", "agentMovementMode": "radial",
                                                             "environmentWidth": 10, "environmentHeight": 8, "startingAgents": 30,
                                                             "agentVision": [1, 6]}),
    # ---- ethics.Asimov (ethics.py:8-98) among plain agents: the three laws over the agents in view.  With spice (trade +
    # tagging + births: children take the class of the parent whose decision model the endowment draw picked), sugar-only
    # with live combat and replacement (humans order robots onto other robots), and on top of diseases / lending / friends
    "asimov_trade": ("trade_tagging_reproduction", 120, dict(_ETHICS, agentDecisionModels=["asimov", "none"])),
    "asimov_combat": ("combat_limited", 120, dict(_ETHICS, agentDecisionModels=["none", "asimov", "asimov"], agentAggressionFactor=[0, 2])),
    # robots next to simple Temperance agents: those answer the second law with their own |delta time to live|
    "asimov_temperance": ("combat_limited", 100, dict(_ETHICS, agentDecisionModels=["asimov", "temperance", "none"], agentAggressionFactor=[0, 2],
                                                      agentDecisionModelFactor=[0.3, 1], agentDynamicDecisionModelFactor=[0.0, 0.1])),
    "asimov_temperance_pecs": ("trade_tagging_reproduction", 100, dict(_ETHICS, agentDecisionModels=["asimov", "temperancePECS", "none", "temperance"],
                                                                       agentDecisionModelFactor=[0.3, 1], agentDynamicSocialPressureFactor=[0.01, 0.1],
                                                                       agentDynamicDecisionModelFactor=[0.0, 0.1])),
    # robots next to the Bentham family: a Bentham neighbour answers the second law with ITS value of the cell, over the
    # neighbourhood it stored at its own last ranking / at birth / at the last (re)placement
    "asimov_bentham": ("trade_tagging_reproduction", 120, dict(_ETHICS, agentDecisionModels=["asimov", "bentham", "none", "egoist", "altruistHalfLookahead"])),
    "asimov_bentham_combat": ("combat_limited", 100, dict(_ETHICS, agentDecisionModels=["asimov", "bentham", "altruist", "none"], agentAggressionFactor=[0, 2],
                                                          agentDecisionModelFactor=[0, 1])),
    # ... with diseases that change vision: the initial population's stored neighbourhoods date from before the first infections
    "asimov_bentham_determinism": ("determinism", 100, dict(_ETHICS, agentDecisionModels=["asimov", "bentham", "egoist", "none"], experimentalGroup=None,
                                                            agentDecisionModelFactor=[0.5, 1], diseaseIncubationPeriod=[0, 0])),
    # (found by fuzzing: radial ranges without wraparound + diseases that change vision + robots next to Bentham agents)
    "asimov_bentham_radial_diseases": ("dune", 40, {
        "seed": 246036, "environmentWidth": 36, "environmentHeight": 36, "startingAgents": 165, "experimentalGroup": None,
        "agentLendingFactor": [0, 0], "agentMaxFriends": [5, 10], "startingDiseases": 30, "startingDiseasesPerAgent": [1, 4],
        "agentImmuneSystemLength": 40, "diseaseTagStringLength": [2, 6], "diseaseSugarMetabolismPenalty": [0, 1], "diseaseVisionPenalty": [-1, 1],
        "diseaseIncubationPeriod": [0, 0], "diseaseTransmissionChance": [0.2, 0.9], "diseaseTimeframe": [0, 100], "agentUniversalSpice": [1, 1],
        "environmentUniversalSugarIncomeInterval": 3, "environmentUniversalSpiceIncomeInterval": 2, "agentDecisionModels": ["asimov", "bentham"],
        "agentDecisionModelFactor": [0.5, 1], "agentDecisionModelLookaheadFactor": 1, "environmentWraparound": False,
        "agentVisionMode": "radial", "agentMovementMode": "radial"}),
    "asimov_determinism": ("determinism", 100, dict(_ETHICS, agentDecisionModels=["asimov", "rawSugarscape"], experimentalGroup=None,
                                                    agentDecisionModelFactor=[0, 1])),
    "ethics_top": ("trade_tagging_reproduction", 100, dict(_ETHICS, agentDecisionModels=["benthamTop", "egoistTop", "altruistNoLookaheadTop", "none", "bentham"])),
    # diffusion timeframe that starts before the first step, delay > 1: the reference's countdown is first decremented
    # at t = 1, so it diffuses at t = 2, 4, ... (the engine fired one step early); and a window
    # that opens later with delay 3
    "pollution_diffusion_delay2": ("pollution_formation", 60, {"environmentPollutionDiffusionTimeframe": [0, 500],
                                                               "environmentPollutionDiffusionDelay": 2,
                                                               "environmentPollutionTimeframe": [1, 500]}),
    "pollution_diffusion_delay3": ("pollution_formation", 60, {"environmentPollutionDiffusionTimeframe": [-1, -1],
                                                               "environmentPollutionDiffusionDelay": 3,
                                                               "environmentPollutionTimeframe": [0, 500]}),
    # other experimental groups than the shipped "depressed" (membership by sex)
    "determinism_group_female": ("determinism", 120, {"experimentalGroup": "female"}),
    "determinism_group_male": ("determinism", 80, {"experimentalGroup": "male"}),
    # ... by race (race<N>: the most common racial tag, fixed at birth) and by decision model (the agent's decisionModel string
    # is the group's name; children carry the string of the parent whose class they take)
    "determinism_group_race1": ("determinism", 100, {"experimentalGroup": "race1"}),
    "determinism_group_bentham": ("determinism", 100, {"experimentalGroup": "bentham"}),
    # ... and a group whose membership CHANGES during the run: the sick (age ranges raise in the reference at the first death:
    # isInGroup reads self.cell.environment of a dead agent)
    "determinism_group_sick": ("determinism", 100, {"experimentalGroup": "sick"}),
    "disease_mild_group_sick": ("disease_basic", 150, dict(_MILD, experimentalGroup="sick")),
    # an experimental group together with agentReplacements: <group>AgentsReplaced / controlAgentsReplaced split the step's
    # replacements by their membership
    "combat_group_male_replacement": ("combat_limited", 100, {"experimentalGroup": "male", "agentAggressionFactor": [0, 2]}),
}


def main():
    names = sys.argv[1:] or sorted(VARIANTS)
    for name in names:
        base, timesteps, overrides = VARIANTS[name]
        path = os.path.join(rh.REFERENCE_DIR, "examples", base + ".json")
        _, meta, steps, log, _ = mg.run_example(path, timesteps, overrides)
        with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt") as f:
            stored = {k: (os.path.basename(v) if k == "environmentFile" else v) for k, v in overrides.items()}   # (relative to this directory)
            json.dump({"meta": meta, "base": base, "overrides": stored, "steps": steps, "log": log}, f)
        sick = [e["sickAgents"] for e in log]
        optimal = [e["agentLastMoveOptimalityPercentage"] for e in log[1:]]
        print(f"{name}: {len(steps) - 1} steps, final pop {steps[-1]['population']}, sick agents min/max {min(sick)}/{max(sick)}, "
              f"move optimality min/max {min(optimal)}/{max(optimal)}", flush=True)


if __name__ == "__main__":
    main()
