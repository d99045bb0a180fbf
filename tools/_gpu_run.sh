mkdir -p gpurun_out
timeout 300 python tools/flow_probe.py --workload S2 --size 4096 --steps 6 --schedulers dataflow,generic,rounds > gpurun_out/r2_t2_probe4096.log 2>&1; echo "probe4096 rc=$?"
tail -c 600 gpurun_out/r2_t2_probe4096.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_t2_launches.csv python tools/flow_probe.py --workload S2 --size 4096 --steps 3 --schedulers dataflow --no-digest > gpurun_out/r2_t2_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python tools/flow_probe.py --workload S2 --size 16384 --steps 5 --schedulers dataflow --no-digest > gpurun_out/r2_t2_probe16384.log 2>&1; echo "probe16384 rc=$?"
tail -c 1500 gpurun_out/r2_t2_probe16384.log
