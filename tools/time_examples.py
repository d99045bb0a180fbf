#!/usr/bin/env python3
"""Wall clock of the shipped examples through the drop-in driver (exact RNG mode): python tools/time_examples.py [names...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import parity_util as pu
    from sugarscape_b200.simulation import Sugarscape
    names = sys.argv[1:] or ["determinism", "all_features", "lending_basic", "disease_basic"]
    out = {}
    for name in names:
        for rep in range(2):                         # (the second run excludes one-off CUDA initialisation)
            c = pu.example_configuration(name, {"timesteps": 200})
            c["logfile"] = None
            S = Sugarscape(c)
            t0 = time.perf_counter()
            S.runSimulation(c["timesteps"])
            dt = time.perf_counter() - t0
        out[name] = {"agent_steps": int(S.agent_steps), "seconds": round(dt, 3), "agent_steps_per_s": round(S.agent_steps / dt, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
