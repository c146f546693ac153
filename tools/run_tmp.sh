timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6
timeout 200 python tools/stream_step_bench.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in d if k in ('value','us_per_step','ms_per_step','unit')})"
timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
