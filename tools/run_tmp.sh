mkdir -p gpurun_out/r2j; O=gpurun_out/r2j
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_decoder" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_bench_shapes.py -q -x -k "not exact_mode and not cfg3" 2>&1 | tail -8
cat gpurun_out/parity_bench_shapes.log
B() { env $2 timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'us/step', d['roofline'].get('us_per_recurrence_step'))"; }
B fused X=1; B fused X=1
B staged CRUSE_FUSE_DECODER=0
timeout 200 python tools/trace_step.py $O/trace_graph_timeline.md --graph > /dev/null 2>$O/trace.err
timeout 200 python tools/wavefront_trace.py $O/wavefront_trace.md > /dev/null 2>&1
