#!/usr/bin/env python3
"""Commit-round profile of one S2 agent step, single GPU or sharded (run under torchrun) -- debug tool."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from sugarscape_b200 import synthetic, steptables, layout, sharded
from sugarscape_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=16384)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--epochs", type=int, default=0)
ap.add_argument("--force-fabric", action="store_true", help="world 1: stitched memory without barriers")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1 or args.force_fabric:
    if world == 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29655")
        dist.init_process_group("gloo", rank=0, world_size=1)
    else:
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
n_agents = int(round(25000000 * (args.size / 16384.0) ** 2))
c = synthetic.configuration(synthetic.S2_OPTIONS, {"environmentWidth": args.size, "environmentHeight": args.size,
                                                   "startingAgents": n_agents, "agentReplacements": n_agents, "timesteps": 1 << 20})
state, params = synthetic.build_state(c)
table = synthetic.endowment_table(state, n_agents)
endow, tags = steptables.step_tables(1, args.steps + 4, c["agentTagStringLength"])
if world > 1 or args.force_fabric:
    gstate, lists = sharded.shard_host_state(state, world)
    eng = sharded.ShardedEngine(params, gstate, lists[rank], rank, world, local, endow, tags)
else:
    eng = Engine(params, state, local, endow, tags)
eng.bind_endowments(table)
if args.epochs > 0:
    eng.set_epochs(args.epochs)
eng.round_trace(1024, fetch=False)
for t in range(1, args.steps + 1):
    eng.step(t, 0.0)
tr = eng.round_trace(1024)
rounds = int(eng.counters()[layout.C_ROUNDS])
tr = tr[:rounds].astype(np.int64)
print(f"rank {rank}/{world} fabric={world > 1 or args.force_fabric} rounds={rounds} list={int(eng.counters()[0])} "
      f"total_us={(tr[-1,5]-tr[0,2])/1e3:.0f} mark {(tr[:,3]-tr[:,2]).sum()/1e3:.0f} select {(tr[:,4]-tr[:,3]).sum()/1e3:.0f} "
      f"exec {(tr[:,5]-tr[:,4]).sum()/1e3:.0f} gap {(tr[1:,2]-tr[:-1,5]).sum()/1e3:.0f} us; active-sum {tr[:,0].sum()}", flush=True)
for r in (0, 1, 2, 50, 100, rounds - 2):
    a = tr[r]
    print(f"  rank {rank} round {r:4d} active {a[0]:7d} ready {a[1]:7d} mark {(a[3]-a[2])/1e3:7.1f} select {(a[4]-a[3])/1e3:7.1f} exec {(a[5]-a[4])/1e3:7.1f}", flush=True)
eng.close()
