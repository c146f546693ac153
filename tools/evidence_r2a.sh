# round-2 first GPU pass (one B200): host topology, full GPU suite (incl. the bench-shape oracle parity tests), smoke, the default
# bench line (with its `train` and `parity` blocks), the reference arm on all 32 clips.  Outputs: gpurun_out/r2a/
mkdir -p gpurun_out/r2a; O=gpurun_out/r2a
( lscpu | head -30; echo; numactl -H 2>&1; echo; nvidia-smi topo -m 2>&1; echo; free -g; echo; nproc; cat /sys/devices/system/node/node*/cpulist 2>&1 ) > $O/host.txt 2>&1
rm -f gpurun_out/parity_bench_shapes.log gpurun_out/grad_parity.log gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; tail -n 25 $O/pytest_gpu.log
cp gpurun_out/parity_bench_shapes.log gpurun_out/grad_parity.log gpurun_out/parity_errors.log $O/ 2>/dev/null
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 600 python bench.py --table $O/kernels_infer.md > $O/bench_infer.json 2>$O/bench_infer.err; tail -n 5 $O/bench_infer.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference.json 2>$O/bench_reference.err
python - <<'PY'
import json
for n in ("infer","reference"):
    try:
        d=json.loads(open(f"gpurun_out/r2a/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms", d.get("ms_per_step"), "value", d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"), d.get("clocks"))
        print("   parity", d.get("parity")); print("   train", {k:v for k,v in (d.get("train") or {}).items() if k!="launch"})
        print("   cpu", d.get("cpu_baseline")); print("   h2d", (d.get("e2e") or {}).get("h2d_alone"), (d.get("e2e") or {}).get("host_binding"))
    except Exception as e: print(n, "ERR", e)
PY
