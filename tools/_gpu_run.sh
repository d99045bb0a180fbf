mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_sharded.py -m gpu -x -q -k "320 or 512 or overrides7 or overrides8" > gpurun_out/r2_d2_sharded.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_d2_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2_d2_bench_n2.json 2> gpurun_out/r2_d2_bench_n2.err; echo "bench2 rc=$?"
tail -c 3000 gpurun_out/r2_d2_bench_n2.json; tail -5 gpurun_out/r2_d2_bench_n2.err
