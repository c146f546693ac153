mkdir -p gpurun_out/r2n
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n/bench_n2.json 2> gpurun_out/r2n/bench_n2.err; tail -n 3 gpurun_out/r2n/bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2n/bench_n2.json").read().strip().splitlines()[-1])
print("N=2 ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("h2d_alone"))
print("train", {k:v for k,v in d["train"].items() if k not in ("launch","workload")})
PY
