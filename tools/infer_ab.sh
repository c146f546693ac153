# A/B of the inference schedule on one GPU: CTA cap of the side kernels (skip convs, decoder groups) beside the recurrences
: > gpurun_out/ms.txt
for cap in 0 64 48 32; do
CRUSE_SIDE_CAP=$cap timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bi_$cap.json 2>gpurun_out/bi_$cap.err
python -c "
import json;d=json.loads(open('gpurun_out/bi_$cap.json').read().strip().splitlines()[-1]);print('side cap=$cap MS',d['ms_per_step'],d['loss'],'e2e',d['e2e']['ms_per_step'])" >> gpurun_out/ms.txt 2>&1 || tail -n 5 gpurun_out/bi_$cap.err >> gpurun_out/ms.txt
done
cat gpurun_out/ms.txt
