mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_decision_models.py -m gpu -x -q -k "determinism_variants" > gpurun_out/r2_j1_moore_pytest.log 2>&1; echo "moore pytest rc=$?"; tail -3 gpurun_out/r2_j1_moore_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reference_trajectory or two_resource or scalable_mode_matches" > gpurun_out/r2_j1_parity_pytest.log 2>&1; echo "parity pytest rc=$?"; tail -3 gpurun_out/r2_j1_parity_pytest.log
