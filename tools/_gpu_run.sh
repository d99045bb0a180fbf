mkdir -p gpurun_out
timeout 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_z5_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_z5_smoke.log
