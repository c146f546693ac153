# A/B of the inference schedule on one GPU: decoder (+ mask*X/iSTFT + loss) pipelined behind layer 2 of the GRU or not
: > gpurun_out/ms.txt
i=0
for cfg in "0 - -" "1 - -" "1 0,3,5,7,8 0,3,5,7,8" "1 0,3,5,6,7,8 0,5,8" "1 0,5,7,8 0,8"; do set -- $cfg
i=$((i+1))
d=$2; k=$3; [ "$d" = "-" ] && d=""; [ "$k" = "-" ] && k=""
CRUSE_PIPELINE_EDGES=$1 CRUSE_DECODE_CUTS=$d CRUSE_SKIP_CUTS=$k timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bi_$i.json 2>gpurun_out/bi_$i.err
python -c "
import json;d=json.loads(open('gpurun_out/bi_$i.json').read().strip().splitlines()[-1]);print('edges=$1 decode=$2 skips=$3 MS',d['ms_per_step'],d['loss'],'e2e',d['e2e']['ms_per_step'])" >> gpurun_out/ms.txt 2>&1 || tail -n 5 gpurun_out/bi_$i.err >> gpurun_out/ms.txt
done
cat gpurun_out/ms.txt
