timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_decoder" 2>&1 | tail -2
python tools/decoder_target.py 55 20 skc; python tools/decoder_target.py 501 20 skc
python tools/decoder_target.py 63 20; python tools/decoder_target.py 501 20
B() { env $2 timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'loss', d['loss'])"; }
B default X=1; B default X=1
