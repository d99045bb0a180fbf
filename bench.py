#!/usr/bin/env python3
"""Benchmark of the per-timestep simulation path (BASELINE.json metric: agent-steps/s and HBM GB/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, see below)

A "step" is one full Sugarscape timestep (environment stencil, agent loop, removal of the dead,
replacement, stats reductions) in scalable RNG mode.  Default workload S2:
16384 x 16384 map, 25 M agents, combat_limited rules -- the same simulation at every N (strong
scaling): one GPU holds it whole, N > 1 shards it by x-slabs over NVLink-stitched memory
(sugarscape_b200/sharded.py), results bit-identical to the single-GPU run.  --workload S1 is the
4096 x 4096 / 1.6 M agents / reproduction_basic shape (one GPU, or N replicas: sharded births are
experimental).

Timed regions
  value   K steps enqueued back to back, state resident in HBM, CUDA events on the launching
          stream, barrier + synchronize on both sides, max over ranks.
  e2e     the same K steps through the public host API (Sugarscape-style doTimestep: stats read
          back to the host every step and turned into the log dict) INCLUDING the upload of the
          whole initial state from host memory and the download of the final agent state.
  roofline  average duration of the dominant kernel (the agent-step kernel) from CUDA events
          recorded inside libssb around each phase, over separately run steps.
  cpu_baseline  the CPU oracle (C port of the reference's algorithm, oracle/ss_oracle.c) on a
          bounded sample of the same workload, one host core.  The reference itself is pure
          Python (about 6-8 k agent-steps/s/core measured in the build container)
          and is not present on the GPU box.
  --impl reference  the same port on ALL host cores: one simulation is strictly serial in the
          reference, its only parallelism is one process per seed (data/run.py:155-180), so the arm
          runs os.cpu_count() independent replicas of a bounded sample (same rules and density).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "agent_steps_per_second"
UNIT = "agent-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=15)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="S2", choices=["S1", "S2"],
                    help="S2 = combat_limited rules with replacement, 16384^2 / 25 M agents (BASELINE configs[4]; the default: "
                         "the same simulation at every N, sharded by x-slabs for N > 1); S1 = reproduction_basic rules, "
                         "4096^2 / 1.6 M agents (configs[3], one GPU or replicas)")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent replicas instead of one sharded simulation")
    ap.add_argument("--size", type=int, default=0, help="map side (S1 = 4096, S2 = 16384)")
    ap.add_argument("--agents", type=int, default=0, help="agents (default: the workload's density: 1.6 M at 4096 / 25 M at 16384)")
    ap.add_argument("--epochs", type=int, default=0)
    ap.add_argument("--mark-shift", type=int, default=-1)
    ap.add_argument("--l2-persist", type=int, default=-1)
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=0, help="timesteps of the CPU port (default: 1 for S2, 3 for S1)")
    return ap.parse_args()


def workload(args, rank=0):
    from sugarscape_b200 import synthetic
    args.cpu_steps = args.cpu_steps or (1 if args.workload == "S2" else 3)
    if args.workload == "S2":
        args.size = args.size or 16384
        n_agents = args.agents or int(round(25000000 * (args.size / 16384.0) ** 2))
        c = synthetic.configuration(synthetic.S2_OPTIONS, {
            "environmentWidth": args.size, "environmentHeight": args.size, "startingAgents": n_agents,
            "agentReplacements": n_agents, "seed": 12345 + rank, "timesteps": 1 << 20})
        return c, n_agents
    args.size = args.size or 4096
    n_agents = args.agents or int(round(1600000 * (args.size / 4096.0) ** 2))
    c = synthetic.configuration(synthetic.S1_OPTIONS, {
        "environmentWidth": args.size, "environmentHeight": args.size, "startingAgents": n_agents,
        "seed": 12345 + rank, "timesteps": 1 << 20})
    return c, n_agents


def describe(args, n_agents):
    rules = "reproduction_basic rules" if args.workload == "S1" else "combat_limited rules (tags, tribes per quadrant, replacement)"
    return f"{args.workload} synthetic {args.size}x{args.size}, {n_agents} agents, {rules}"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [p.strip() for p in out.stdout.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# algorithmic bytes (canonical dtypes): growback 12 B per cell; agent step with
# reproduction rules 232 B per acting agent
BYTES_PER_CELL_GROWBACK = 12
BYTES_PER_AGENT_STEP = {"S1": 232, "S2": 196}


def aggregate_over_ranks(elapsed, units, device=None):
    """Whole-job numbers from per-rank ones: time = MAX over ranks (the slowest rank defines the
    step), work = SUM over ranks.  Works with any initialised torch.distributed backend."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elapsed), float(units)
    t = torch.tensor([float(elapsed)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])


def run_cpu_port(args, c, steps, state=None, table=None):
    """The oracle port on one host core over `steps` timesteps of the same initial state
    (`state`, if given, is consumed)."""
    from oracle_binding import Oracle
    from sugarscape_b200 import synthetic, steptables, layout, init
    if state is None:
        state, params = synthetic.build_state(c)
    else:
        params = layout.make_params(c, layout.RNG_SCALABLE, init.rule_flags(c), init.max_agent_range(c), init.equator(c))
    endow, tags = steptables.step_tables(1, steps + 2, c["agentTagStringLength"])
    orc = Oracle(params, state, endow, tags)
    replacing = c["agentReplacements"] > 0
    if replacing:
        orc.bind_endowments(table if table is not None else synthetic.endowment_table(state, int(c["startingAgents"])))
    agent_steps = 0
    t0 = time.perf_counter()
    for t in range(1, steps + 1):
        agent_steps += int(state.counters[0])
        orc.env_step(t)
        orc.agent_step(t, 0.0)
        orc.compact(t)
        if replacing:
            orc.replace(t)
        orc.stats(t)
    dt = time.perf_counter() - t0
    return agent_steps / dt, dt, agent_steps


def _cpu_replica(job):
    """One worker of the all-core CPU arm: an independent replica (own seed) of a bounded sample."""
    options, overrides, steps = job
    from sugarscape_b200 import synthetic
    c = synthetic.configuration(options, overrides)
    value, dt, agent_steps = run_cpu_port(None, c, steps)
    return agent_steps, dt


def run_cpu_port_all_cores(args, c, steps):
    """The reference's own parallelism is one process per seed (reference data/run.py:155-180): one
    simulation cannot use a second core.  All-core CPU figure = os.cpu_count() independent replicas
    of a bounded sample of the workload (same rules and density on a map of 1/64 the area; the cost
    per agent-step is flat in the map size), each timed around its own step loop."""
    import multiprocessing as mp
    from sugarscape_b200 import synthetic
    cores = os.cpu_count() or 1
    side = max(256, args.size // 8)
    n = max(1, int(round(c["startingAgents"] * (side / args.size) ** 2)))
    options = synthetic.S2_OPTIONS if args.workload == "S2" else synthetic.S1_OPTIONS
    jobs = []
    for k in range(cores):
        ov = {"environmentWidth": side, "environmentHeight": side, "startingAgents": n, "seed": 12345 + k, "timesteps": 1 << 20}
        if c["agentReplacements"] > 0:
            ov["agentReplacements"] = n
        jobs.append((options, ov, steps))
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_replica, jobs)
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return total / slowest, slowest, total, cores, f"{cores} replicas of {side}x{side} / {n} agents, {steps} timesteps each"


class Job:
    """Host state of the workload (built once per process) and engines over it."""

    def __init__(self, args, c, n_agents, rank, world, local_rank):
        from sugarscape_b200 import sharded, steptables, synthetic
        self.args, self.c, self.n_agents, self.rank, self.world, self.local_rank = args, c, n_agents, rank, world, local_rank
        self.sharded = world > 1 and not args.replicas
        self.endow, self.tags = steptables.step_tables(1, args.warmup + 4 * args.steps + 64, c["agentTagStringLength"])
        self.state, self.params = synthetic.build_state(c)
        self.table = synthetic.endowment_table(self.state, n_agents) if c["agentReplacements"] > 0 else None
        self.lists, self.engines = None, 0
        if self.sharded:
            self.state, self.lists = sharded.shard_host_state(self.state, world)

    def engine(self):
        from sugarscape_b200 import sharded
        from sugarscape_b200.engine import Engine
        a = self.args
        self.engines += 1
        if self.sharded:
            eng = sharded.ShardedEngine(self.params, self.state, self.lists[self.rank], self.rank, self.world, self.local_rank,
                                        self.endow, self.tags, mark_shift=a.mark_shift if a.mark_shift >= 0 else None,
                                        tag=str(self.engines))
            h2d = sum((eng.cell_hi - eng.cell_lo if not k.startswith("a_") else eng.slot_hi - eng.slot_lo) * v.element_size()
                      for k, v in eng.tensors.items() if isinstance(v, sharded.StitchedArray))
            h2d += sum(v.numel() * v.element_size() for v in eng.tensors.values() if not isinstance(v, sharded.StitchedArray))
        else:
            eng = Engine(self.params, self.state, self.local_rank, self.endow, self.tags)
            if a.mark_shift >= 0:
                eng.set_mark_shift(a.mark_shift)
            h2d = sum(v.numel() * v.element_size() for v in eng.tensors.values())
        if self.table is not None:
            eng.bind_endowments(self.table)
            h2d += sum(v.numel() * v.element_size() for v in eng.endowment_tensors.values())
        if a.epochs > 0:
            eng.set_epochs(a.epochs)
        if a.l2_persist >= 0:
            eng.set_l2_persist(bool(a.l2_persist))
        return eng, h2d

    def check(self, eng):
        from sugarscape_b200 import layout
        code = int(eng.counters()[layout.C_OVERFLOW])
        if code != 0 or (self.sharded and eng.broken()):
            raise RuntimeError(f"rank {self.rank}: engine failure flag {code} (1 = agent capacity, 7 = box-wide barrier lost)")


def device_resident_leg(job, eng, args, barrier, dev):
    """K steps enqueued back to back on resident state; returns the figures of this rank."""
    import torch
    t = 0
    for _ in range(args.warmup):
        t += 1
        eng.step(t, 0.0)
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(job.local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    first = t + 1
    for _ in range(args.steps):
        t += 1
        eng.step(t, 0.0)
    ev1.record()
    barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    # agents acting in step s = population after step s - 1 (of this rank's list when sharded)
    start_pop = len(job.lists[job.rank]) if job.sharded else job.n_agents
    acting = [int(eng.stats_history(s - 1).population) if s > 1 else start_pop for s in range(first, t + 1)]
    rounds = [int(eng.stats_history(s).rounds) for s in range(first, t + 1)]
    job.check(eng)
    # per-phase kernel times (separate steps, CUDA events inside libssb)
    eng.enable_timing(True)
    phase = {"env_ms": [], "agent_ms": [], "compact_ms": [], "stats_ms": []}
    phase_acting = []
    for _ in range(min(args.steps, 10)):
        phase_acting.append(int(eng.stats_history(t).population))
        t += 1
        eng.step(t, 0.0)
        tim = eng.timing()
        for k in phase:
            phase[k].append(tim[k])
    eng.enable_timing(False)
    barrier()
    return {"ms": ms, "launches": launches, "acting": acting, "rounds": rounds, "clocks": sampler.summary(),
            "phase": {k: float(np.mean(v)) for k, v in phase.items()}, "phase_acting": float(np.mean(phase_acting))}


def e2e_leg(job, args, barrier, dev):
    """The same K steps through the host API: upload of the initial state, per-step stats readback
    turned into the log dict, download of the final agents."""
    from sugarscape_b200 import init, layout, stats as stats_mod
    barrier()
    w0 = time.perf_counter()
    eng, h2d = job.engine()                                          # H2D of the whole initial state
    sb = stats_mod.StatsBuilder(job.c, init.equator(job.c))
    get_stats = eng.global_stats if job.sharded else eng.stats
    eng.reduce_stats(0)
    rs = sb.update(get_stats(), 0)
    agent_steps = 0
    for s in range(1, args.steps + 1):
        agent_steps += rs["population"]
        eng.step(s, rs["meanWealth"])
        rs = sb.update(get_stats(), s)                              # D2H of the step's reductions
        json.dumps(rs)                                              # the log line the CLI writes
    n = int(eng.counters()[layout.C_N_LIST])
    order = eng.tensors["a_order"][:n].cpu().numpy()                # D2H of this rank's final agents
    d2h = order.nbytes + args.steps * C.sizeof(layout.Stats)
    if job.sharded:
        lo, hi = eng.slot_lo, eng.slot_hi
        for k in ("a_id", "a_pos", "a_sugar", "a_age"):
            d2h += eng.tensors[k][lo:hi].cpu().numpy().nbytes
    else:
        final = eng.download(["a_id", "a_pos", "a_sugar", "a_age", "counters"])
        d2h += sum(getattr(final, k).nbytes for k in ("a_id", "a_pos", "a_sugar", "a_age", "counters"))
    barrier()
    wall = time.perf_counter() - w0
    job.check(eng)
    eng.close()
    if job.sharded:                       # every rank counted the box-wide population
        agent_steps = agent_steps / job.world
    return wall, agent_steps, h2d, d2h


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "S1" and not os.environ.get("SSB_EXPERIMENTAL_SHARDED_BIRTHS"):
        args.replicas = True            # births are not sharded yet (experimental): S1 runs on one GPU or as replicas
    sharded_run = world > 1 and not args.replicas

    if args.impl == "reference":
        if rank != 0:
            return 0
        c, n_agents = workload(args)
        steps = max(3, min(args.steps, 8))
        value, dt, agent_steps, cores, sample = run_cpu_port_all_cores(args, c, steps)
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": 0, "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "S2" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": describe(args, n_agents), "rng_mode": "scalable"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample + "; oracle/ss_oracle.c (C port of the reference's algorithm), one process per "
                                       "core like the reference's own data/run.py pool -- one simulation is strictly serial "
                                       "(the reference itself is pure Python and absent from this box: 6-8 k agent-steps/s/core "
                                       "measured in the build container, BASELINE.md)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # sharded: ONE simulation, the same seed on every rank; replicas: one simulation per rank
    c, n_agents = workload(args, 0 if sharded_run else rank)
    job = Job(args, c, n_agents, rank, world, local_rank)
    n_cells = args.size * args.size

    eng, _ = job.engine()
    res = device_resident_leg(job, eng, args, barrier, dev)
    eng.close()
    del eng
    torch.cuda.empty_cache()
    ms, agent_steps = aggregate_over_ranks(res["ms"], sum(res["acting"]), dev)
    value = agent_steps / (ms / 1000.0)
    launches = res["launches"]
    avg = res["phase"]
    peak, peak_kind = measured_peak_gbs()
    agent_bytes = res["phase_acting"] * BYTES_PER_AGENT_STEP[args.workload]
    achieved = agent_bytes / (avg["agent_ms"] / 1000.0) / 1e9
    env_bytes = (n_cells // world if sharded_run else n_cells) * BYTES_PER_CELL_GROWBACK
    roofline = {"bound": "hbm", "kernel": "k_agent_step_scalable", "achieved": achieved, "peak": peak,
                "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6.65 TB/s",
                "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "traffic_note": "not captured for this launch; ncu --set full of the same kernel on the 4096^2 map with S2 rules: "
                                "8.9 GB DRAM traffic per launch = 5.7 KB per agent-step, 29x the algorithmic bytes "
                                "(profiles/r1_agent_step_s2rules_4096_ncu_full.csv)",
                "per": "rank 0's GPU",
                "algorithmic_bytes_per_launch": agent_bytes, "avg_launch_ms": avg["agent_ms"],
                "env_step": {"kernel": "k_growback_v4", "achieved": env_bytes / (avg["env_ms"] / 1000.0) / 1e9,
                             "frac": env_bytes / (avg["env_ms"] / 1000.0) / 1e9 / peak, "avg_launch_ms": avg["env_ms"],
                             "note": "12 B/cell assumed; the store is skipped for vectors already at capacity, so > 1 is possible"
                                     + ("; includes the box-wide barrier that follows the stencil" if sharded_run else "")},
                "phase_ms": avg}

    e2e = None
    if not args.skip_e2e:
        wall, e2e_steps, h2d, d2h = e2e_leg(job, args, barrier, dev)
        wall, e2e_steps = aggregate_over_ranks(wall, e2e_steps, dev)
        e2e = {"value": e2e_steps / wall, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps,
               "d2h_bytes_per_step": d2h / args.steps, "wall_s": wall,
               "includes": "upload of the initial state, per-step stats readback + log formatting, download of the final agents"
                           + (" (bytes per rank)" if world > 1 else "")}

    # ---- CPU baseline (rank 0, N = 1 only; runs last: it consumes the host state) ----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        v, dt, n_as = run_cpu_port(args, c, args.cpu_steps, job.state, job.table)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{args.cpu_steps} timesteps of the same {args.workload} initial state ({n_as} agent-steps, {dt:.1f} s), "
                         "oracle/ss_oracle.c single thread; the Python reference itself runs 6-8 k agent-steps/s/core "
                         "(build container, BASELINE.md) and cannot instantiate this map"}

    if rank == 0:
        if sharded_run:
            parallelism = f"x-slabs over {world} GPUs (one simulation; NVLink-stitched state, box-wide commit rounds)"
        else:
            parallelism = "replicas" if world > 1 else "single GPU"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if (sharded_run or world == 1) and args.workload == "S2" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": describe(args, n_agents) + (f", {world} independent replicas" if world > 1 and not sharded_run else ""),
                       "rng_mode": "scalable", "l2": "inputs larger than L2 (state > 500 MB per GPU)",
                       "parallelism": parallelism,
                       "mean_commit_rounds": float(np.mean(res["rounds"])), "mean_acting_agents_rank0": float(np.mean(res["acting"]))},
            "clocks": res["clocks"], "gpu_launches": launches, "roofline": roofline,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
