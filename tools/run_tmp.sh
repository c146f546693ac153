mkdir -p gpurun_out/r3
timeout 300 python -m pytest tests/test_gpu_bwd.py tests/test_gpu_bench_shapes.py -q -x -m gpu -k "captured_train or cfg3 or train_step" 2>&1 | tail -2
timeout 200 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r3/bt.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms', round(d['ms_per_step'],4), 'loss', d['loss'], 'e2e', round(d['e2e']['ms_per_step'],4))"
timeout 100 python tools/trace_step.py gpurun_out/r3/trace_train_timeline.md --graph --train > /dev/null 2>gpurun_out/r3/trace.err; tail -4 gpurun_out/r3/trace_train_timeline.md | cut -c1-110
