mkdir -p gpurun_out/r3
timeout 900 python -m pytest tests/test_gpu_bwd.py tests/test_gpu_bench_shapes.py -q -x -m gpu 2>&1 | tail -3
python tools/determinism_check.py 2>&1 | tail -8 | tee gpurun_out/r3/determinism.log
T() { env $2 timeout 200 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r3/bt_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms', round(d['ms_per_step'],4), 'loss', d['loss'], 'e2e', round(d['e2e']['ms_per_step'],4))"; }
T new "X=1"
