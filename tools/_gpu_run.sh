mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scalable_mode or full_size" > gpurun_out/r2_m1_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_m1_pytest.log
for size in 4096 8192 16384; do
SSB_SWEEP_DIAG=1 timeout 600 python tools/flow_probe.py --workload S2 --size $size --steps 5 --schedulers sweep --no-digest > gpurun_out/r2_m1_probe${size}.log 2>&1
python - <<PY
import json
l=json.loads(open("gpurun_out/r2_m1_probe${size}.log").read().strip().splitlines()[0])
print($size, round(l["mean_agent_ms"],2), {k:v for k,v in l["steps"][-1].items() if k.startswith("agent_")}, l["sweep_diag"])
PY
done
