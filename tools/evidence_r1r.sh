# round-1 (r1r) evidence pass on one B200: full GPU suite, smoke, both bench workloads (+ tables), reference arm, streaming
# step, kernel timeline, ncu launch list (summaries land in gpurun_out/r1r/, copied to profiles/ by hand)
mkdir -p gpurun_out/r1r; O=gpurun_out/r1r
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -n 3 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 400 python bench.py --table $O/kernels_infer.md > $O/bench_infer.json 2>$O/bench_infer.err; tail -c 600 $O/bench_infer.json | head -c 300; echo
timeout 400 python bench.py --workload train --no-cpu-baseline --table $O/kernels_train.md > $O/bench_train.json 2>$O/bench_train.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>$O/bench_reference.err
timeout 200 python tools/stream_step_bench.py --table $O/kernels_stream.md > $O/bench_stream.json 2>$O/stream.err
timeout 200 python tools/trace_step.py $O/trace_graph_timeline.md --graph > /dev/null 2>$O/trace.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
python - <<'PY'
import json
for n in ("infer","train","reference","stream"):
    try:
        d=json.loads(open(f"gpurun_out/r1r/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("ms_per_step", d.get("us_per_step")), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
    except Exception as e: print(n, "ERR", e)
PY
