mkdir -p gpurun_out/r2n
( nvidia-smi topo -m; lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket"; free -g | head -2 ) > gpurun_out/r2n/host_n8.txt 2>&1
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2n/bench_n$n.json 2> gpurun_out/r2n/bench_n$n.err; tail -n 2 gpurun_out/r2n/bench_n$n.err
done
python - <<'PY'
import json
for n in (8,4):
    d=json.loads(open(f"gpurun_out/r2n/bench_n{n}.json").read().strip().splitlines()[-1])
    print(f"N={n} ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "e2e ms", d["e2e"]["ms_per_step"], d["e2e"].get("h2d_alone"), d["e2e"].get("host_binding"))
    print("train", {k:v for k,v in d["train"].items() if k not in ("launch","workload")})
PY
cat gpurun_out/r2n/host_n8.txt | head -30
