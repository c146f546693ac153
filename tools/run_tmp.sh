timeout 600 python -m pytest tests/test_gpu_frontend.py -q -x 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>gpurun_out/b.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'pcm16', d['e2e'].get('pcm16'))" || tail -5 gpurun_out/b.err
