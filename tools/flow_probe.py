#!/usr/bin/env python3
"""A/B probe of the two mode S schedulers on a synthetic workload (one GPU).

    python tools/flow_probe.py --workload S2 --size 8192 --steps 6 [--schedulers sweep,sweep-generic,dataflow,rounds]

Runs the same initial state through each scheduler, prints the per-phase CUDA-event times of every step and
compares the ID-sorted state digests after the last step (they must be equal: both reproduce the sequential
priority order).  Output: one JSON line per scheduler + one comparison line.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="S2")
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--schedulers", default="sweep,rounds")
    ap.add_argument("--no-digest", action="store_true")
    args = ap.parse_args()
    import parity_util as pu
    from sugarscape_b200 import synthetic, steptables, layout
    from sugarscape_b200.engine import Engine
    if args.workload == "S2":
        n = int(round(25000000 * (args.size / 16384.0) ** 2))
        c = synthetic.configuration(synthetic.S2_OPTIONS, {"environmentWidth": args.size, "environmentHeight": args.size,
                                                           "startingAgents": n, "agentReplacements": n, "timesteps": 1 << 20})
    else:
        n = int(round(1600000 * (args.size / 4096.0) ** 2))
        c = synthetic.configuration(synthetic.S1_OPTIONS, {"environmentWidth": args.size, "environmentHeight": args.size,
                                                           "startingAgents": n, "timesteps": 1 << 20})
    state, params = synthetic.build_state(c)
    endow, tags = steptables.step_tables(1, args.steps + 8, c["agentTagStringLength"])
    table = synthetic.endowment_table(state, n) if c["agentReplacements"] > 0 else None
    digests = {}
    for sched in args.schedulers.split(","):
        eng = Engine(params, state, 0, endow, tags)
        if table is not None:
            eng.bind_endowments(table)
        eng.set_scheduler("rounds" if sched.startswith("rounds") else sched)
        if sched.endswith("ms2-e256"):
            eng.set_mark_shift(2)
            eng.set_epochs(256)
        eng.enable_timing(True)
        rows = []
        for t in range(1, args.steps + 1):
            eng.step(t, 0.0)
            tim = eng.timing()
            st = eng.stats()
            rows.append({"t": t, "population": int(st.population), "rounds": int(st.rounds), **{k: round(v, 3) for k, v in tim.items()}})
        code = int(eng.counters()[layout.C_OVERFLOW])
        diag = eng.sweep_diag() if hasattr(eng, "sweep_diag") else None
        line = {"scheduler": sched, "workload": args.workload, "size": args.size, "agents": n, "overflow": code, "steps": rows,
                "mean_agent_ms": float(np.mean([r["agent_ms"] for r in rows[1:]] or [rows[0]["agent_ms"]]))}
        if diag:
            line["sweep_diag"] = {"warp_iterations": diag[0], "held": diag[1], "ran": diag[2], "behind_watermark": diag[3]}
        if not args.no_digest:
            d, snap = pu.state_digests_by_id(eng.download())
            digests[sched] = d
            line["digest"] = {k: v[:16] for k, v in d.items()}
        print(json.dumps(line), flush=True)
        eng.close()
        del eng
    if len(digests) > 1:
        names = list(digests)
        same = all(digests[names[0]] == digests[k] for k in names[1:])
        print(json.dumps({"schedulers_agree": same, "compared": names}), flush=True)
        return 0 if same else 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
