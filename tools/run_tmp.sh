mkdir -p gpurun_out/r2f
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3
for i in 1 2; do timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"; done
timeout 200 python tools/trace_step.py gpurun_out/r2f/trace_graph_timeline.md --graph > /dev/null 2>&1
timeout 200 python tools/wavefront_trace.py gpurun_out/r2f/wavefront_trace.md 2>&1 | grep SUMMARY
