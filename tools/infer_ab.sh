# A/B of the inference schedule on one GPU: encoder head / decoder tail pipelined with the GRU wavefront or not
: > gpurun_out/ms.txt
for e in 0 1; do
CRUSE_PIPELINE_EDGES=$e timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bi_$e.json 2>gpurun_out/bi_$e.err
python -c "
import json;d=json.loads(open('gpurun_out/bi_$e.json').read().strip().splitlines()[-1]);print('edges=$e MS',d['ms_per_step'],d['loss'],'e2e',d['e2e']['ms_per_step'])" >> gpurun_out/ms.txt 2>&1 || tail -n 5 gpurun_out/bi_$e.err >> gpurun_out/ms.txt
done
cat gpurun_out/ms.txt
