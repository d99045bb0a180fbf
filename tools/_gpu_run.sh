mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_decision_models.py -m gpu -x -q -k "nowrap or determinism_variants" > gpurun_out/r2_i1_nowrap_pytest.log 2>&1; echo "nowrap pytest rc=$?"; tail -3 gpurun_out/r2_i1_nowrap_pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --launch-skip 3200 --launch-count 260 --csv --log-file gpurun_out/r2_i1_steady_kernels.csv python tools/flow_probe.py --workload S2 --size 16384 --steps 36 --schedulers sweep --no-digest > gpurun_out/r2_i1_ncu.log 2>&1; echo "ncu rc=$?"
tail -c 600 gpurun_out/r2_i1_ncu.log
