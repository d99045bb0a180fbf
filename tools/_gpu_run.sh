mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/r2_n1_bench_n8.json 2> gpurun_out/r2_n1_bench_n8.err; echo "bench8 rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/r2_n1_bench_n8.json").read().strip().splitlines()[-1])
print(8, l["value"], l["ms_per_step"], l["state_digest"]["value"], {k: round(v,2) for k,v in l["roofline"]["phase_ms"].items()})
PY
tail -2 gpurun_out/r2_n1_bench_n8.err
