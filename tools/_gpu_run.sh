mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scalable_mode and (sweep or patched)" > gpurun_out/r2_t7_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_t7_pytest.log
timeout 300 python tools/flow_probe.py --workload S2 --size 4096 --steps 6 --schedulers sweep,rounds > gpurun_out/r2_t7_probe4096.log 2>&1; echo "probe4096 rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2_t7_probe4096.log"):
    try: d = json.loads(l)
    except Exception: continue
    if "scheduler" in d: print(d["scheduler"], d["mean_agent_ms"], d["overflow"])
    else: print(d)
PY
timeout 600 python tools/flow_probe.py --workload S2 --size 16384 --steps 5 --schedulers sweep --no-digest > gpurun_out/r2_t7_probe16384.log 2>&1; echo "probe16384 rc=$?"
tail -c 700 gpurun_out/r2_t7_probe16384.log
timeout 1500 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:"k_sweep_lean" -s 1 -c 1 -f -o gpurun_out/r2_t7_sweep16384 python tools/flow_probe.py --workload S2 --size 16384 --steps 2 --schedulers sweep --no-digest > gpurun_out/r2_t7_ncu.log 2>&1; echo "ncu rc=$?"
