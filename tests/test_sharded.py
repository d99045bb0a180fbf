"""x-slab sharding (sugarscape_b200/sharded.py): host logic on CPU, box-wide parity on >= 2 GPUs."""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import parity_util as pu
from sugarscape_b200 import layout, sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_and_slot_bounds_partition_everything():
    for width, world in ((50, 2), (51, 4), (16384, 8), (7, 8)):
        b = sharded.slab_bounds(width, world)
        assert b[0] == 0 and b[-1] == width and all(b[i] <= b[i + 1] for i in range(world))
    assert sharded.slot_bounds(1000, 3) == [0, 333, 666, 1000]


def test_shard_host_state_moves_every_record_to_the_rank_that_owns_its_cell():
    c, state, pop, params, endow, tags = pu.build("combat_limited", layout.RNG_SCALABLE, {"agentMovement": [6, 6]})
    world = 4
    new, lists = sharded.shard_host_state(state, world)
    xb, sb = new.slab_bounds, sharded.slot_bounds(state.capacity, world)
    assert xb[0] == 0 and xb[-1] == state.width and all(xb[i] < xb[i + 1] for i in range(world))
    sizes = [len(x) for x in lists]
    assert max(sizes) - min(sizes) <= 0.1 * sum(sizes) / world + state.height, sizes      # balanced by population
    assert sum(len(x) for x in lists) == int(state.counters[layout.C_N_LIST])
    for r, slots in enumerate(lists):
        assert np.all((slots >= sb[r]) & (slots < sb[r + 1]))
        x = new.a_pos[slots] // state.height
        assert np.all((x >= xb[r]) & (x < xb[r + 1]))
        assert np.array_equal(new.occ[new.a_pos[slots]], slots)
    # the same simulation: identical ID-sorted snapshot
    new.a_order[:] = -1
    order = np.concatenate(lists)
    new.a_order[:len(order)] = order
    d_old, _ = pu.state_digests_by_id(state)
    d_new, _ = pu.state_digests_by_id(new)
    assert d_old == d_new


def test_slabs_of_whole_tiles_for_the_sharded_sweep():
    """The ranks can run the sweep scheduler over their slabs when every slab consists of whole
    64-column tiles and the rule set is the plain lean transaction's: wide maps get such slabs from the balancer, the
    eligibility test mirrors the engine's (csrc/ssb_abi.cu, shard_sweep_possible)."""
    wide = {"environmentWidth": 512, "environmentHeight": 150, "startingAgents": 12000, "agentMovement": [6, 6]}
    c, state, pop, params, endow, tags = pu.build("constant_growback", layout.RNG_SCALABLE, wide)
    for world in (2, 4):
        new, lists = sharded.shard_host_state(state, world)
        assert all(b % sharded.SWEEP_TILE == 0 for b in new.slab_bounds[:-1]), new.slab_bounds
        assert sharded.sweep_slabs_possible(params, new.slab_bounds)
        sizes = [len(x) for x in lists]
        assert max(sizes) - min(sizes) <= 0.25 * sum(sizes) / world, sizes          # still balanced by population
    assert not sharded.sweep_slabs_possible(params, [0, 100, 512])                  # a slab boundary inside a tile
    # rule sets whose transaction reads a neighbour's record or dilates its footprint stay on the reservation rounds
    for name, overrides in (("combat_unlimited", {"agentAggressionFactor": [1, 1]}), ("trade_basic", {}), ("pollution_formation", {}),
                            ("cultural_tagging", {})):
        c2, s2, pop2, params2, e2, t2 = pu.build(name, layout.RNG_SCALABLE, dict(wide, environmentHeight=512, **overrides))
        assert not sharded.sweep_slabs_possible(params2, [0, 256, 512]), name
    # narrow maps keep the page-aligned / plain balancing (no whole tiles to give)
    c3, s3, pop3, params3, e3, t3 = pu.build("constant_growback", layout.RNG_SCALABLE, {"agentMovement": [6, 6]})
    new3, _ = sharded.shard_host_state(s3, 2)
    assert not sharded.sweep_slabs_possible(params3, new3.slab_bounds)


def test_chunk_boundaries_are_page_multiples_and_follow_the_shares():
    g = sharded.GRANULE
    b = sharded.chunk_boundaries([100, 200, 300, 400], 4, 4)            # tiny array: one page per rank
    assert b == [0, g, 2 * g, 3 * g, 4 * g]
    ends = [3 * 1024 * 1024, 4 * 1024 * 1024, 9 * 1024 * 1024, 16 * 1024 * 1024]      # elements; 4-byte items
    b = sharded.chunk_boundaries(ends, 4, 4)
    assert b[0] == 0 and b[-1] >= ends[-1] * 4 and all(x % g == 0 for x in b) and all(b[i] < b[i + 1] for i in range(4))
    assert [x // 4 for x in b[1:]] == ends


def test_combine_stats_sums_counts_and_keeps_extrema():
    a, b = layout.Stats(), layout.Stats()
    a.timestep = b.timestep = 7
    a.population, b.population = 10, 0
    a.sum_wealth, b.sum_wealth = 55.5, 0.0
    a.max_wealth, a.min_wealth = 9.0, 1.0
    b.max_wealth, b.min_wealth = -1e300, 1e300        # an empty rank must not leak its sentinels
    a.tribe_count[1], b.tribe_count[1] = 4, 6
    a.tribe_first[1], b.tribe_first[1] = 3, -1
    a.lorenz_weighted_sum = b.lorenz_weighted_sum = 123.0
    a.rounds, b.rounds = 140, 140
    out = sharded.combine_stats([a, b])
    assert (out.timestep, out.population, out.sum_wealth, out.max_wealth, out.min_wealth) == (7, 10, 55.5, 9.0, 1.0)
    assert out.tribe_count[1] == 10 and out.tribe_first[1] == 3 and out.lorenz_weighted_sum == 123.0 and out.rounds == 140


_FD_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from sugarscape_b200 import sharded
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = []
for i in range(150):         # more than one SCM_RIGHTS message
    r, w = os.pipe()
    os.write(w, b"hello %d from %d" % (i, rank))
    mine.append(r)
directory = sharded.rendezvous_directory(rank, dist)       # rank 0's private directory, announced through the group
assert os.path.isdir(directory) and (os.stat(directory).st_mode & 0o077) == 0
fds = sharded.exchange_fds(rank, world, mine, "t" + os.environ["MASTER_PORT"], dist.barrier, directory)
for src, got in sorted(fds.items()):
    assert len(got) == 150
    if src != rank:          # (the peer drains this rank's pipes)
        for i, fd in enumerate(got):
            assert os.read(fd, 64) == b"hello %d from %d" % (i, src), (rank, src, i)
print("FD_EXCHANGE_OK", rank, flush=True)
dist.destroy_process_group()
"""


def test_fd_exchange_between_two_gloo_ranks():
    """The arena descriptors travel between the ranks over unix sockets: checked with pipes."""
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(_FD_WORKER)
    try:
        res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                              "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), f.name, ROOT],
                             capture_output=True, text=True, timeout=120)
    finally:
        os.unlink(f.name)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("FD_EXCHANGE_OK") == 2


SHARDED_CASES = [
    ("constant_growback", {}, 40),
    ("combat_unlimited", {"agentAggressionFactor": [1, 1]}, 40),
    ("cultural_tagging", {"agentFertilityFactor": [0, 0]}, 40),      # tagging across the slab boundary
    ("trade_basic", {}, 40),
    ("pollution_formation", {}, 60),
    ("combat_limited", {}, 60),
    ("combat_limited", {"environmentWidth": 300, "environmentHeight": 300, "startingAgents": 4000, "agentReplacements": 4000,
                        "agentMaxAge": [5, 30], "agentAggressionFactor": [0, 1]}, 40),
    # maps wide enough for slabs of whole 64-column tiles and rule sets of the plain lean transaction: the ranks run the SWEEP
    # scheduler over their slabs (done bits, cells and watermarks across NVLink) -- the S2 configuration of bench.py
    ("combat_limited", {"environmentWidth": 320, "environmentHeight": 200, "startingAgents": 9000, "agentReplacements": 9000,
                        "agentMaxAge": [5, 30]}, 40),
    ("constant_growback", {"environmentWidth": 512, "environmentHeight": 150, "startingAgents": 12000, "agentVision": [1, 6]}, 30),
]


# Births across slabs: passed on one 2-GPU box, diverged at t = 25 of cultural_tagging on another.
# The engine refuses the combination unless SSB_EXPERIMENTAL_SHARDED_BIRTHS is set; so do these cases.
SHARDED_BIRTH_CASES = [
    ("cultural_tagging", {}, 40),
    ("reproduction_basic", {"environmentWidth": 100, "environmentHeight": 100, "startingAgents": 1500,
                            "environmentSugarPeaks": [[70, 30, 4], [30, 70, 4], [20, 20, 4], [80, 80, 4]]}, 50),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,overrides,steps", SHARDED_BIRTH_CASES)
def test_two_gpu_slabs_with_births_experimental(name, overrides, steps):
    if not os.environ.get("SSB_EXPERIMENTAL_SHARDED_BIRTHS"):
        pytest.skip("sharded births are experimental (set SSB_EXPERIMENTAL_SHARDED_BIRTHS=1)")
    test_slabs_equal_the_single_gpu_run(name, overrides, steps, 2)


# The world sizes a box offers are compared with the single-GPU run: 2 GPUs on a 2-GPU lease, 2 and 4 on a larger box (the
# 50-wide examples give slabs of 12 columns at world 4 -- about an agent's reach, so a transaction can touch three
# slabs).  A 1-GPU lease skips them (the CPU suite covers the host side of the sharding, above).
@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("name,overrides,steps", SHARDED_CASES)
def test_slabs_equal_the_single_gpu_run(name, overrides, steps, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    if world == 4 and overrides.get("environmentWidth") == 512 and not os.environ.get("SSB_TEST_ALL_WORLDS"):
        # round 2 ran out of 4-GPU time before this one case (the other eight passed at world 4, and bench.py's S2 run has the
        # same state checksum at 1 / 2 / 4 / 8 GPUs); it must not be the first thing a 4-GPU box ever runs under `pytest -x`
        pytest.skip("not yet run at world 4 (SSB_TEST_ALL_WORLDS=1 to include it)")
    res = subprocess.run(["timeout", "300", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29900 + os.getpid() % 90),
                          os.path.join(ROOT, "tests", "sharded_worker.py"), "--case", name, "--overrides", json.dumps(overrides),
                          "--steps", str(steps)], capture_output=True, text=True, timeout=360)
    log = res.stdout + res.stderr
    errors = [ln for ln in log.splitlines() if "Error" in ln or "rror:" in ln or "SystemExit" in ln or "differs" in ln]
    assert res.returncode == 0, "\n".join(errors[:12]) or log[-3000:]
    assert res.stdout.count("SHARDED_PARITY_OK") == world, res.stdout[-2000:]
    # every rank runs the same scheduler; the last two cases must run the sweep when the map gives every rank slabs of whole
    # 64-column tiles whatever the population balance
    counts = {k: res.stdout.count(f"scheduler {k} ") for k in ("sweep", "rounds")}
    assert sorted(counts.values()) == [0, world], f"the ranks disagree on the scheduler: {counts}"
    if SHARDED_CASES.index((name, overrides, steps)) >= len(SHARDED_CASES) - 2 and 128 * world <= overrides["environmentWidth"]:
        assert counts["sweep"] == world, "expected the sweep scheduler:\n" + res.stdout[-1500:]
