#!/usr/bin/env python3
"""Per-agent log fixtures from the UNMODIFIED reference (build container only): the reference's own
runSimulation() with `agentLogfile` set, on small maps so that the fixtures stay small.

    python tests/golden/make_golden_agentlog.py

Writes agentlog.json.gz = {case: {"base", "overrides", "agent_log": [...], "log": [...]}}."""
import gzip
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

SMALL = {"environmentWidth": 20, "environmentHeight": 20, "startingAgents": 70}
CASES = {
    # every rule family at once + the decision models + the experimental group
    "determinism": ("determinism", dict(SMALL, timesteps=30)),
    # live combat (preyKilled / preyWealth), tribes
    "combat": ("cultural_combat_unlimited", dict(SMALL, timesteps=30, agentAggressionFactor=[0, 2])),
    # pollution (pollutionDifference), trade partners, births
    "trade_pollution": ("trade_pollution", dict(SMALL, timesteps=30, agentFertilityFactor=[1, 1], agentFemaleFertilityAge=[12, 15],
                                                agentMaleFertilityAge=[12, 15], agentFemaleInfertilityAge=[40, 50],
                                                agentMaleInfertilityAge=[50, 60], agentMaxAge=[60, 100])),
    # radial ranges + ethics.Asimov among plain agents (a robot that stays although it has candidates logs moveRank = validMoves)
    "radial_asimov": ("determinism", dict(SMALL, timesteps=30, agentVisionMode="radial", agentMovementMode="radial",
                                          agentDecisionModels=["asimov", "none"], agentDecisionModelFactor=[0, 1], experimentalGroup=None,
                                          diseaseVisionPenalty=[-1, 0], diseaseMovementPenalty=[-1, 0])),
}


def main():
    out = {}
    for name, (base, ov) in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            ov = dict(ov, agentLogfile=os.path.join(tmp, "agents.json"))
            S = rh.build_reference(os.path.join(rh.REFERENCE_DIR, "examples", base + ".json"), ov, logfile=os.path.join(tmp, "log.json"))
            S.runSimulation(S.configuration["timesteps"])
            agent_log = json.load(open(ov["agentLogfile"]))
            log = json.load(open(os.path.join(tmp, "log.json")))
        stored = {k: v for k, v in ov.items() if k != "agentLogfile"}
        out[name] = {"base": base, "overrides": stored, "agent_log": agent_log, "log": log}
        print(f"{name}: {len(agent_log)} agent records over {len(log) - 1} steps, "
              f"{sum(r['preyKilled'] for r in agent_log)} kills, {sum(r['mates'] for r in agent_log)} matings", flush=True)
    with gzip.open(os.path.join(HERE, "agentlog.json.gz"), "wt") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
