#!/usr/bin/env python3
"""Prints the commit-round profile of one mode S agent step of the S1 workload (debug tool)."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sugarscape_b200 import synthetic, steptables, layout
from sugarscape_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--steps", type=int, default=25)
ap.add_argument("--epochs", type=int, default=0)
ap.add_argument("--mark-shift", type=int, default=-1)
ap.add_argument("--l2-persist", type=int, default=-1)
args = ap.parse_args()
n_agents = int(round(1600000 * (args.size / 4096.0) ** 2))
c = synthetic.configuration(synthetic.S1_OPTIONS, {"environmentWidth": args.size, "environmentHeight": args.size,
                                                   "startingAgents": n_agents, "timesteps": 1 << 20})
state, params = synthetic.build_state(c)
endow, tags = steptables.step_tables(1, args.steps + 4, 0)
eng = Engine(params, state, 0, endow, tags)
if args.epochs > 0:
    eng.set_epochs(args.epochs)
if args.mark_shift >= 0:
    eng.set_mark_shift(args.mark_shift)
if args.l2_persist >= 0:
    eng.set_l2_persist(bool(args.l2_persist))
eng.round_trace(1024, fetch=False)
for t in range(1, args.steps + 1):
    eng.step(t, 0.0)
    if t in (2, 10, args.steps):
        tr = eng.round_trace(1024)
        rounds = int(eng.counters()[layout.C_ROUNDS])
        tr = tr[:rounds].astype(np.int64)
        print(f"t={t} rounds={rounds} population={int(eng.counters()[0])} total_us={(tr[-1,5]-tr[0,2])/1e3:.0f}")
        print("  round  active   ready  mark_us select_us exec_us")
        for r in list(range(0, min(rounds, 12))) + list(range(12, rounds, max(1, rounds // 12))):
            a = tr[r]
            print(f"  {r:5d} {a[0]:7d} {a[1]:7d} {(a[3]-a[2])/1e3:8.1f} {(a[4]-a[3])/1e3:8.1f} {(a[5]-a[4])/1e3:8.1f}")
        print("  sums: mark %.0f us, select %.0f us, exec %.0f us, active-sum %d" % (
            (tr[:,3]-tr[:,2]).sum()/1e3, (tr[:,4]-tr[:,3]).sum()/1e3, (tr[:,5]-tr[:,4]).sum()/1e3, tr[:,0].sum()))
