# A/B of the inference schedule on one GPU: length of the last wavefront chunk (frames; 0 = equal chunks)
: > gpurun_out/ms.txt
for v in 0 32 16 0 32; do
CRUSE_LAST_CHUNK=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bi_$v.json 2>gpurun_out/bi_$v.err
python -c "
import json;d=json.loads(open('gpurun_out/bi_$v.json').read().strip().splitlines()[-1]);print('last chunk=$v MS',d['ms_per_step'],d['loss'],'e2e',d['e2e']['ms_per_step'])" >> gpurun_out/ms.txt 2>&1 || tail -n 5 gpurun_out/bi_$v.err >> gpurun_out/ms.txt
done
cat gpurun_out/ms.txt
