timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_decoder" 2>&1 | tail -3
python tools/decoder_target.py 63; python tools/decoder_target.py 125; python tools/decoder_target.py 501 
B() { env $2 $3 timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"; }
B default X=1
B spare24 CRUSE_SIDE_SPARE=24
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -x 2>&1 | tail -3
