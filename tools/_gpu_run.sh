mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scalable_mode and (sweep or patched)" > gpurun_out/r2_c3_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_c3_pytest.log
for v in "" _np; do
SSB_LIBSSB=sugarscape_b200/libssb$v.so SSB_SWEEP_DIAG=1 timeout 600 python tools/flow_probe.py --workload S2 --size 16384 --steps 4 --schedulers sweep --no-digest > gpurun_out/r2_c3_probe16384$v.log 2>&1; echo "probe16384$v rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/r2_c3_probe16384$v.log").read().strip().splitlines()[-1])
print(l["steps"][-1], l["sweep_diag"])
PY
done
timeout 600 python tools/flow_probe.py --workload S2 --size 4096 --steps 6 --schedulers sweep,rounds > gpurun_out/r2_c3_probe4096.log 2>&1; echo "probe4096 rc=$?"
tail -c 200 gpurun_out/r2_c3_probe4096.log
