# A/B of the inference schedule on one GPU: number of flag chunks of the wavefront
: > gpurun_out/ms.txt
for v in 8 6 10 12 8; do
CRUSE_FLAG_CHUNKS=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bi_$v.json 2>gpurun_out/bi_$v.err
python -c "
import json;d=json.loads(open('gpurun_out/bi_$v.json').read().strip().splitlines()[-1]);print('flag chunks=$v MS',d['ms_per_step'],d['loss'],'e2e',d['e2e']['ms_per_step'])" >> gpurun_out/ms.txt 2>&1 || tail -n 5 gpurun_out/bi_$v.err >> gpurun_out/ms.txt
done
cat gpurun_out/ms.txt
