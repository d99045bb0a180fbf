mkdir -p gpurun_out
timeout 36 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29933 tests/sharded_worker.py --case constant_growback --overrides '{}' --steps 8 > gpurun_out/r2_z3_sharded_min.log 2>&1; echo "worker rc=$?"; grep -c SHARDED_PARITY_OK gpurun_out/r2_z3_sharded_min.log; tail -2 gpurun_out/r2_z3_sharded_min.log | cut -c1-300
