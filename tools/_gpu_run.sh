mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2_d5_bench_n$n.json 2> gpurun_out/r2_d5_bench_n$n.err; echo "bench$n rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/r2_d5_bench_n$n.json").read().strip().splitlines()[-1])
print($n, l["value"], l["ms_per_step"], l["state_digest"]["value"], {k: round(v,2) for k,v in l["roofline"]["phase_ms"].items()}, l["e2e"]["value"], l["e2e"]["split_rank0"])
PY
tail -2 gpurun_out/r2_d5_bench_n$n.err
done
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --skip-cpu --skip-extra > gpurun_out/r2_d5_bench_n1.json 2> gpurun_out/r2_d5_bench_n1.err; echo "bench1 rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/r2_d5_bench_n1.json").read().strip().splitlines()[-1])
print(1, l["value"], l["ms_per_step"], l["state_digest"]["value"], {k: round(v,2) for k,v in l["roofline"]["phase_ms"].items()}, l["e2e"]["value"])
PY
