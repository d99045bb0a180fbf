mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scalable_mode and (sweep or patched)" > gpurun_out/r2_t12_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_t12_pytest.log
timeout 600 python tools/flow_probe.py --workload S2 --size 16384 --steps 4 --schedulers sweep --no-digest > gpurun_out/r2_t12_probe16384.log 2>&1; echo "probe16384 rc=$?"
tail -c 600 gpurun_out/r2_t12_probe16384.log
timeout 1500 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section MemoryWorkloadAnalysis --section Occupancy --clock-control none --import-source on -k regex:"k_sweep_lean" -s 1 -c 1 -f -o gpurun_out/r2_t12_sweep16384 python tools/flow_probe.py --workload S2 --size 16384 --steps 2 --schedulers sweep --no-digest > gpurun_out/r2_t12_ncu.log 2>&1; echo "ncu rc=$?"
