timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3
for i in 1 2; do timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"; done
for e in "CRUSE_SIDE_SPARE=20" "CRUSE_SIDE_SPARE=8"; do env $e timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$e ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"; done
