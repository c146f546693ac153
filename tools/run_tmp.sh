B() { env $2 $3 $4 timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'loss', d['loss'])"; }
B default X=1
B default X=1
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_bench_shapes.py -q -x -k "not exact_mode and not cfg3" 2>&1 | tail -4
mkdir -p gpurun_out/r2l
timeout 200 python tools/trace_step.py gpurun_out/r2l/trace_graph_timeline.md --graph > /dev/null 2>&1
