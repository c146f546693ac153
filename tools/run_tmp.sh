B() { timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'loss', d['loss'])"; }
echo prefetch; python tools/conv_target.py 10; B prefetch
cp cruse_b200/libcruse_sm100.so /tmp/def.so; cp variants/lib_nopf.so cruse_b200/libcruse_sm100.so
echo no-prefetch; python tools/conv_target.py 10; B noprefetch
cp /tmp/def.so cruse_b200/libcruse_sm100.so
B prefetch
timeout 300 python bench.py --workload train --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', round(d['ms_per_step'],4))"
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -x 2>&1 | tail -2
