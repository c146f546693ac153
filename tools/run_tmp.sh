B() { env $2 $3 $4 timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'loss', d['loss'])"; }
B tail_plain X=1
B tail_skc CRUSE_TAIL_PLAIN_DECODER=0
B tail_plain X=1
B tail_skc CRUSE_TAIL_PLAIN_DECODER=0
B tail_plain_last48 CRUSE_LAST_CHUNK=48
B tail_plain_last32 CRUSE_LAST_CHUNK=32
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_bench_shapes.py -q -x -k "not exact_mode and not cfg3" 2>&1 | tail -4
cat gpurun_out/parity_bench_shapes.log | head -2
mkdir -p gpurun_out/r2o
timeout 200 python tools/trace_step.py gpurun_out/r2o/trace_graph_timeline.md --graph > /dev/null 2>&1
