#!/usr/bin/env python3
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py [example ...]

For every example config (reference examples/*.json) under the reference's own test protocol
(tests/test.py:17-43: headlessMode, debugMode none, timesteps 200) it records

  <name>.json.gz : meta, per-step digests of the canonical state (ref_harness.digests),
                   per-step (population, sum of agent IDs), and the reference's log entries
                   (runtimeStats after every step, exactly what log.json holds)
  <name>.npz     : full canonical snapshots at SNAP_STEPS (t=0 is the state after init)

The fixtures pin the oracle (oracle/) and, through it and directly, the CUDA engine.
"""
import glob
import gzip
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

SNAP_STEPS = [0, 1, 2, 3, 5, 10, 20, 50, 100, 150, 200]
TIMESTEPS = 200


def run_example(path, timesteps=TIMESTEPS, overrides=None):
    name = os.path.splitext(os.path.basename(path))[0]
    ov = {"timesteps": timesteps}
    ov.update(overrides or {})
    S = rh.build_reference(path, ov)
    S.updateRuntimeStats()  # runSimulation does this before the loop (sugarscape.py:863)
    steps = []
    snaps = {}
    log = [dict(S.runtimeStats)]
    snap = rh.snapshot(S)
    snaps[0] = snap
    steps.append({"t": 0, "population": len(S.agents), "id_sum": int(snap["a_id"].sum()),
                  "digests": rh.digests(snap)})
    t0 = time.time()
    agent_steps = 0
    for t in range(1, timesteps + 1):
        if len(S.agents) == 0 and not S.keepAlive:
            break
        agent_steps += len(S.agents)
        S.doTimestep()
        snap = rh.snapshot(S)
        steps.append({"t": S.timestep, "population": len(S.agents),
                      "id_sum": int(snap["a_id"].sum()), "digests": rh.digests(snap)})
        log.append(dict(S.runtimeStats))
        if t in SNAP_STEPS:
            snaps[t] = snap
    wall = time.time() - t0
    meta = {"example": name, "timesteps": timesteps, "seed": S.seed,
            "width": S.environment.width, "height": S.environment.height,
            "reference_agent_steps": agent_steps, "reference_loop_seconds": round(wall, 3),
            "python": sys.version.split()[0]}
    return name, meta, steps, log, snaps


def save(outdir, name, meta, steps, log, snaps):
    with gzip.open(os.path.join(outdir, name + ".json.gz"), "wt") as f:
        json.dump({"meta": meta, "steps": steps, "log": log}, f)
    flat = {}
    for t, snap in snaps.items():
        for k, v in snap.items():
            if k in ("sugar", "spice") and np.all(v == np.floor(v)):
                v = v.astype(np.int32)
            elif k in ("occ", "a_id", "a_x", "a_y", "a_age", "a_tags", "a_tribe", "a_alive"):
                v = v.astype(np.int32)
            flat[f"t{t}_{k}"] = v
    np.savez_compressed(os.path.join(outdir, name + ".npz"), **flat)


def dump_example_options(paths):
    """The options each shipped example sets (its JSON minus the __README__ text), as ONE
    fixture file; the tests materialise temporary --conf files from it."""
    out = {}
    for p in paths:
        with open(p) as f:
            d = json.load(f)
        d = d.get("sugarscapeOptions", d)
        out[os.path.splitext(os.path.basename(p))[0]] = {k: v for k, v in d.items() if not k.startswith("__")}
    with open(os.path.join(HERE, "example_configs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def main():
    names = sys.argv[1:]
    ex_dir = os.path.join(rh.REFERENCE_DIR, "examples")
    paths = sorted(glob.glob(os.path.join(ex_dir, "*.json")))
    dump_example_options(paths)
    if names:
        paths = [p for p in paths if os.path.splitext(os.path.basename(p))[0] in names]
    for p in paths:
        t0 = time.time()
        name, meta, steps, log, snaps = run_example(p)
        save(HERE, name, meta, steps, log, snaps)
        print(f"{name}: {len(steps) - 1} steps, final pop {steps[-1]['population']}, "
              f"{meta['reference_agent_steps']} agent-steps, "
              f"{meta['reference_agent_steps'] / max(meta['reference_loop_seconds'], 1e-9):.0f}/s, "
              f"{time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
