mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -m gpu -x -q --durations=8 ) > gpurun_out/r2_e1_pytest_full.log 2>&1; echo "pytest rc=$?"
tail -22 gpurun_out/r2_e1_pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_e1_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_e1_smoke.log
