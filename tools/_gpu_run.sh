mkdir -p gpurun_out
SSB_SWEEP_DIAG=1 timeout 600 python tools/flow_probe.py --workload S2 --size 16384 --steps 4 --schedulers sweep --no-digest > gpurun_out/r2_h1_probe16384.log 2>&1; echo "probe16384 rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/r2_h1_probe16384.log").read().strip().splitlines()[0])
print(l["mean_agent_ms"], {k:v for k,v in l["steps"][-1].items() if k.startswith("agent")}, l["sweep_diag"])
PY
( time timeout 2400 python -m pytest tests/ -m gpu -x -q ) > gpurun_out/r2_h1_pytest_full.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_h1_pytest_full.log
timeout 1200 python bench.py > gpurun_out/r2_h1_bench.json 2> gpurun_out/r2_h1_bench.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/r2_h1_bench.json").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["steps"], l["parity_at_scale"], l["roofline"]["frac"], {k: round(v,2) for k,v in l["roofline"]["phase_ms"].items()}, l["e2e"], l.get("s1",{}).get("value"), {k:round(v.get("value",0)) for k,v in l["e_configs"].items()})
PY
tail -3 gpurun_out/r2_h1_bench.err
