# A/B on one GPU: programmatic dependent launch of the conv stages on / off (inference and training step)
: > gpurun_out/ms.txt
for v in 1 0 1 0; do
CRUSE_CONV_PDL=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bi_$v.json 2>gpurun_out/bi_$v.err
python -c "
import json;d=json.loads(open('gpurun_out/bi_$v.json').read().strip().splitlines()[-1]);print('infer pdl=$v MS',d['ms_per_step'],d['loss'],'e2e',d['e2e']['ms_per_step'])" >> gpurun_out/ms.txt 2>&1 || tail -n 5 gpurun_out/bi_$v.err >> gpurun_out/ms.txt
done
for v in 1 0; do
CRUSE_CONV_PDL=$v timeout 300 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bt_$v.json 2>gpurun_out/bt_$v.err
python -c "
import json;d=json.loads(open('gpurun_out/bt_$v.json').read().strip().splitlines()[-1]);print('train pdl=$v MS',d['ms_per_step'],d['loss'],'e2e',d['e2e']['ms_per_step'])" >> gpurun_out/ms.txt 2>&1 || tail -n 5 gpurun_out/bt_$v.err >> gpurun_out/ms.txt
done
cat gpurun_out/ms.txt
