"""The N > 1 plumbing of bench.py on CPU: two gloo ranks, whole-job time = max over ranks,
whole-job work = sum over ranks (each rank runs an independent replica of the workload)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
import bench
dist.init_process_group("gloo")
rank = dist.get_rank()
elapsed, units = bench.aggregate_over_ranks(10.0 + 5.0 * rank, 1000.0 * (rank + 1))
assert elapsed == 15.0, elapsed
assert units == 3000.0, units
# the reference arm: only rank 0 works, the others exit 0 without output
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank%d-ok\n" % rank)
"""


def test_aggregate_over_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 150), str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank0-ok" in out.stdout and "rank1-ok" in out.stdout


def test_reference_arm_only_rank0_prints(tmp_path):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--size", "256", "--cpu-steps", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--size", "256", "--cpu-steps", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and '"impl": "reference"' in out.stdout
    # the contract of the reference arm's line
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["metric"] == "agent_steps_per_second" and line["unit"] == "agent-steps/s" and line["higher_is_better"] is True
    assert "workload" in line["config"] and "model" not in line["config"] and "ran" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == line["value"] and cb["sample"] and cb["state_after"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["vs_baseline"] is None and line["value"] > 0
