mkdir -p gpurun_out/r2g
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
for e in "CRUSE_HEAD_CHUNKS=0" "" "CRUSE_HEAD_CHUNKS=2" "CRUSE_HEAD_CHUNKS=4" "CRUSE_HEAD_CHUNKS=3 CRUSE_HEAD_CAP=100" "CRUSE_HEAD_CHUNKS=3 CRUSE_HEAD_CAP=64" "CRUSE_HEAD_CHUNKS=2 CRUSE_HEAD_CAP=100" "CRUSE_HEAD_CHUNKS=1 CRUSE_HEAD_CAP=116"; do
 for i in 1 2; do env $e timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$e', 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['loss'])"; done
done
timeout 200 python tools/wavefront_trace.py gpurun_out/r2g/wavefront_trace.md 2>&1 | grep -v "^$" | tail -22
