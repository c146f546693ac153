# round-2 multi-GPU pass on one 8-GPU box (gpurun --gpus 8): host topology, the H2D probe (pinned vs write-combined staging, all ranks
# at once), the default bench line at N = 8 (pinned and --wc staging), N = 4, N = 2.  Outputs: gpurun_out/r2n/
mkdir -p gpurun_out/r2n; O=gpurun_out/r2n
( nvidia-smi topo -m; lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket"; free -g | head -2 ) > $O/host_n8.txt 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 tools/h2d_probe.py > $O/h2d_probe_n8.txt 2>$O/h2d_probe.err; cat $O/h2d_probe_n8.txt
run() { n=$1; tag=$2; shift 2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 "$@" > $O/bench_n$n$tag.json 2> $O/bench_n$n$tag.err; grep -E "NCCL communicator|Init COMPLETE" $O/bench_n$n$tag.err | head -2 | cut -c1-200
python - <<PY
import json
d=json.loads(open("$O/bench_n$n$tag.json").read().strip().splitlines()[-1])
print("N=$n$tag ms", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],4), "pcm16", round((d["e2e"].get("pcm16") or {}).get("value", 0)), round((d["e2e"].get("pcm16") or {}).get("ms_per_step", 0), 4), d["e2e"].get("staging"), {k: round(v,1) if isinstance(v,float) else v for k,v in d["e2e"]["h2d_alone"].items() if k!="note"})
t=d.get("train") or {}
print("   train", {k:v for k,v in t.items() if k not in ("launch","workload","allreduce")}, (t.get("allreduce") or {}).get("us"), (t.get("allreduce") or {}).get("bus_GBps"), (t.get("allreduce") or {}).get("pct_of_step"))
PY
}
run 8 ""
run 4 ""
run 2 ""
