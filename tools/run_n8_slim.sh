# the default bench line at N = 8 only (final build): gpurun --gpus 8 -- 'bash tools/run_n8_slim.sh'.  Output: gpurun_out/r2n/bench_n8.json
mkdir -p gpurun_out/r2n; O=gpurun_out/r2n
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; grep -E "NCCL communicator" $O/bench_n8.err | head -1 | cut -c1-200
python - <<PY
import json
d=json.loads(open("gpurun_out/r2n/bench_n8.json").read().strip().splitlines()[-1])
print("N=8 ms", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],4), "pcm16", round((d["e2e"].get("pcm16") or {}).get("value", 0)), {k: round(v,1) if isinstance(v,float) else v for k,v in d["e2e"]["h2d_alone"].items() if k!="note"})
t=d.get("train") or {}
print("   train", {k:v for k,v in t.items() if k not in ("launch","workload","allreduce")}, (t.get("allreduce") or {}).get("us"), (t.get("allreduce") or {}).get("bus_GBps"), (t.get("allreduce") or {}).get("pct_of_step"))
PY
