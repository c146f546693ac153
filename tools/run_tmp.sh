timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_decoder" 2>&1 | tail -2
B() { env $2 $3 $4 timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"; }
B default X=1
B cap72 CRUSE_DEC_CAP=72
B cap84 CRUSE_DEC_CAP=84
B cap48 CRUSE_DEC_CAP=48
B cuts_single CRUSE_DECODE_CUTS=0,1,2,3,4,5,6,7,8
B cuts_0357 CRUSE_DECODE_CUTS=0,3,5,7,8
B cuts_2468 CRUSE_DECODE_CUTS=0,2,4,6,8
B spare32 CRUSE_SIDE_SPARE=32
B default X=1
