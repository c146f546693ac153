mkdir -p gpurun_out/r3
timeout 900 python -m pytest tests/test_gpu_bwd.py tests/test_gpu_bench_shapes.py -q -x -m gpu 2>&1 | tail -2
T() { env $2 timeout 200 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r3/bt_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms', round(d['ms_per_step'],4), 'loss', d['loss'], 'e2e', round(d['e2e']['ms_per_step'],4))"; }
T cap0 "CRUSE_BWD_TAIL_CAP=0"


T cap0 "CRUSE_BWD_TAIL_CAP=0"
timeout 200 python tools/trace_step.py gpurun_out/r3/trace_train_timeline.md --graph --train > /dev/null 2>gpurun_out/r3/trace.err; grep -c colsum gpurun_out/r3/trace_train_timeline.md; grep "colsum" gpurun_out/r3/trace_train_timeline.md | awk -F'|' '{s+=$3} END {print "colsum total us", s}'
