B() { timeout 200 python bench.py --no-cpu-baseline --no-train-block --steps 30 $2 2>gpurun_out/wc.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['e2e']['staging'], d['e2e']['h2d_alone']['GBps_per_rank'])" || tail -5 gpurun_out/wc.err; }
B pinned
B wc --wc
python tools/h2d_probe.py
