#!/usr/bin/env python3
"""Long-run fixtures from the UNMODIFIED reference: the examples whose kin lists keep growing (reproduction with
inheritance / trade), run for their SHIPPED number of timesteps (or 1000), with the state digests of every 50th
step.  They pin what the 200-step fixtures cannot: pool recycling (the kin pool's entries of removed agents are
reused) and anything else that only shows after hundreds of generations.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_long.py

Writes tests/golden/long_runs.json.gz."""
import gzip
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

CASES = {"trade_reproduction": 1000, "reproduction_inheritance": 1000}
EVERY = 50


def run(name, timesteps):
    S = rh.build_reference(os.path.join(rh.REFERENCE_DIR, "examples", name + ".json"), {"timesteps": timesteps})
    S.updateRuntimeStats()
    steps = []
    t0 = time.time()
    for t in range(1, timesteps + 1):
        if len(S.agents) == 0 and not S.keepAlive:
            break
        S.doTimestep()
        if t % EVERY == 0 or t == timesteps:
            snap = rh.snapshot(S)
            steps.append({"t": t, "population": len(S.agents), "digests": rh.digests(snap)})
    print(f"{name}: {timesteps} steps in {time.time() - t0:.1f} s, final population {len(S.agents)}", flush=True)
    return {"timesteps": timesteps, "steps": steps}


def main():
    out = {name: run(name, n) for name, n in CASES.items()}
    with gzip.open(os.path.join(HERE, "long_runs.json.gz"), "wt") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
