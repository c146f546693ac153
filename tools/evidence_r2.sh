# round-2 evidence pass on one B200 (outputs: gpurun_out/r2b/, copied to profiles/*_r2b* by hand; r2 = the mid-round pass, r2b = the final build): full GPU suite, smoke, the default bench
# line (with `train` and `parity` blocks) + per-kernel table, the train workload, the reference arm on all 32 clips, the streaming step,
# kernel timeline, wavefront progress trace, the ncu launch list of the bench command and full-set captures of the main kernels
mkdir -p gpurun_out/r2b; O=gpurun_out/r2b
rm -f gpurun_out/parity_bench_shapes.log gpurun_out/grad_parity.log gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; tail -n 3 $O/pytest_gpu.log
cp gpurun_out/parity_bench_shapes.log gpurun_out/grad_parity.log gpurun_out/parity_errors.log $O/ 2>/dev/null
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -n 1 $O/smoke.log
timeout 600 python bench.py --table $O/kernels_infer.md > $O/bench_infer.json 2>$O/bench_infer.err; tail -n 3 $O/bench_infer.err
timeout 600 python bench.py --workload train --no-cpu-baseline --table $O/kernels_train.md > $O/bench_train.json 2>$O/bench_train.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference.json 2>$O/bench_reference.err
timeout 200 python tools/stream_step_bench.py --table $O/kernels_stream.md > $O/bench_stream.json 2>$O/stream.err
timeout 200 python tools/trace_step.py $O/trace_graph_timeline.md --graph > /dev/null 2>$O/trace.err
timeout 200 python tools/wavefront_trace.py $O/wavefront_trace.md > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-block > $O/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gru_seq_tc_kernel|gemm_astat_tc_kernel' -c 4 -o $O/ncu_gru_full python tools/ncu_target.py gru > $O/ncu_gru.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel|stft512|mask_istft512|wo_male_partial|layernorm|enc1_stream|dec1_stream' -c 24 -o $O/ncu_side_full python tools/ncu_target.py side > $O/ncu_side.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decoder_fused_kernel -c 2 --launch-skip 3 -o $O/ncu_decoder_full python tools/decoder_target.py 501 3 skc > $O/ncu_decoder.log 2>&1
timeout 200 python tools/trace_step.py $O/trace_train_timeline.md --graph --train > /dev/null 2>>$O/trace.err
ls -la $O | head -40
python - <<'PY'
import json
for n in ("infer","train","reference","stream"):
    try:
        d=json.loads(open(f"gpurun_out/r2b/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("ms_per_step", d.get("us_per_step")), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
    except Exception as e: print(n, "ERR", e)
PY
