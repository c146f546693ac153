# round-2 closing evidence pass on one B200 (r2c = the build with the re-scheduled training step and the MN-major GEMM operands; the
# inference-side kernels are unchanged since r2b, whose ncu captures / launch list / wavefront trace stand).  Outputs: gpurun_out/r2c/
# (copied to profiles/*_r2c* by hand).  ncu reports are summarised ON THE BOX and deleted: gpurun brings back at most 64 MiB.
mkdir -p gpurun_out/r2c; O=gpurun_out/r2c
rm -f gpurun_out/parity_bench_shapes.log gpurun_out/grad_parity.log gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; tail -n 3 $O/pytest_gpu.log
cp gpurun_out/parity_bench_shapes.log gpurun_out/grad_parity.log gpurun_out/parity_errors.log $O/ 2>/dev/null
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -n 1 $O/smoke.log
timeout 600 python bench.py --table $O/kernels_infer.md > $O/bench_infer.json 2>$O/bench_infer.err; tail -n 3 $O/bench_infer.err
timeout 600 python bench.py --workload train --no-cpu-baseline --table $O/kernels_train.md > $O/bench_train.json 2>$O/bench_train.err
timeout 200 python tools/trace_step.py $O/trace_train_timeline.md --graph --train > /dev/null 2>$O/trace.err
timeout 100 python tools/gemm_mn_check.py > $O/gemm_operand_major_check.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'gemm_tn_tc_kernel|gru_bwd_tc_kernel' -c 10 -o $O/ncu_train_full python bench.py --workload train --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_train.log 2>&1
python tools/ncu_summary.py $O/ncu_train_full.ncu-rep > $O/ncu_train_full.md 2>>$O/ncu_train.log; rm -f $O/*.ncu-rep
du -sh gpurun_out; ls -la $O | head -30
python - <<'PY'
import json
for n in ("infer","train"):
    try:
        d=json.loads(open(f"gpurun_out/r2c/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("ms_per_step", d.get("us_per_step")), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
        if n == "infer": print("  train block", {k: v for k, v in (d.get("train") or {}).items() if k != "launch"}); print("  parity", d.get("parity"))
    except Exception as e: print(n, "ERR", e)
PY
