mkdir -p gpurun_out/r3
T() { timeout 200 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r3/bt_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms', round(d['ms_per_step'],4), 'loss', d['loss'], 'e2e', round(d['e2e']['ms_per_step'],4))"; }
P() { rm -f gpurun_out/parity_bench_shapes.log; timeout 600 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_bwd.py -q -m gpu -k "cfg3 or train or full_model or grouped_gru" 2>&1 | tail -2; grep -i "worst" gpurun_out/parity_bench_shapes.log | cut -c1-260; }
echo "== default"; T default; P
cp cruse_b200/libcruse_sm100.so /tmp/keep.so; cp variants/lib_tanhtrain.so cruse_b200/libcruse_sm100.so
echo "== tanh gates in the training forward"; T tanh; P
cp /tmp/keep.so cruse_b200/libcruse_sm100.so
