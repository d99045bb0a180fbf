"""Worker of tests/test_sharded.py: run under torchrun with one rank per GPU.

Every rank first runs the whole simulation on its own GPU (single-GPU engine, scalable mode) and
records the per-step digests of the ID-sorted state; then the ranks run the same simulation
sharded by x-slabs and compare the box-wide state with those digests after every step.
"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import parity_util as pu  # noqa: E402
from sugarscape_b200 import layout, sharded  # noqa: E402
from sugarscape_b200.engine import Engine  # noqa: E402

KEYS = ("agents_int", "agents_f64", "agents_tags", "grid_int", "pollution")
SUMS = ("population", "n_born", "sum_sugar_metab", "sum_vision", "sum_age", "n_dead", "dead_starvation", "dead_aging", "dead_combat",
        "n_replaced", "env_wealth_total", "sum_neighbors", "sum_valid_moves")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True)
    ap.add_argument("--overrides", default="{}")
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    overrides = dict({"agentMovement": [6, 6]}, **json.loads(args.overrides))
    c, state, pop, params, endow, tags = pu.build(args.case, layout.RNG_SCALABLE, overrides)
    replacing = c["agentReplacements"] > 0

    single = Engine(params, state.copy(), local, endow, tags)
    if replacing:
        single.bind_endowments(pop.endowments)
    expected = []
    for t in range(1, args.steps + 1):
        single.step(t, 0.0)
        d, snap = pu.state_digests_by_id(single.download())
        expected.append((d, len(snap["a_id"]), single.stats()))
    assert int(single.counters()[layout.C_OVERFLOW]) == 0
    single.close()

    global_state, lists = sharded.shard_host_state(state, world)
    eng = sharded.ShardedEngine(params, global_state, lists[rank], rank, world, local, endow, tags)
    if replacing:
        eng.bind_endowments(pop.endowments)
    remote, migrated = 0, 0
    for t in range(1, args.steps + 1):
        eng.step(t, 0.0)
        if eng.broken():
            raise SystemExit(f"rank {rank}: box-wide barrier broke at t={t}")
        assert int(eng.counters()[layout.C_OVERFLOW]) == 0, f"rank {rank}: engine failure flag {int(eng.counters()[layout.C_OVERFLOW])} at t={t}"
        hs = eng.download_global()
        d, snap = pu.state_digests_by_id(hs)
        want, want_pop, want_stats = expected[t - 1]
        assert len(snap["a_id"]) == want_pop, f"t={t}: population {len(snap['a_id'])} != {want_pop}"
        for k in KEYS:
            assert d[k] == want[k], f"rank {rank}: {k} differs from the single-GPU run at t={t}"
        assert int(eng.counters()[layout.C_OVERFLOW]) == 0
        st = eng.global_stats()
        for f in SUMS:
            assert getattr(st, f) == getattr(want_stats, f), f"stats.{f} at t={t}: {getattr(st, f)} != {getattr(want_stats, f)}"
        for f in ("sum_wealth", "lorenz_weighted_sum", "max_wealth", "min_wealth", "sum_happiness"):
            a, b = float(getattr(st, f)), float(getattr(want_stats, f))
            assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), f"stats.{f} at t={t}: {a} != {b}"
        assert list(st.tribe_count) == list(want_stats.tribe_count)
        n = int(eng.counters()[layout.C_N_LIST])
        pos = hs.a_pos[eng.tensors["a_order"][:n].cpu().numpy()]
        remote = int(np.sum((pos < eng.cell_lo) | (pos >= eng.cell_hi)))
        assert remote == 0, f"rank {rank}: {remote} agents of its list stand on another rank's slab after the migration at t={t}"
        slots = eng.tensors["a_order"][:n].cpu().numpy()
        if not (params.flags & layout.F_REPRO):     # (a child's record is allocated by the acting parent's rank)
            assert np.all((slots >= eng.slot_lo) & (slots < eng.slot_hi)), "a record outside this rank's slot range"
        migrated += int(eng.counters()[layout.C_N_MIGRATED])
    print(f"SHARDED_PARITY_OK rank {rank}/{world} case {args.case} steps {args.steps} scheduler {eng.scheduler()} "
          f"population {want_pop} migrated away {migrated}", flush=True)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
