# A/B of the training-step schedule on one GPU (captured graph): side-stream work on / off, repeated to see the run-to-run spread
: > gpurun_out/ms.txt
for cfg in "1 0" "1 1" "1 0" "1 1" "0 0"; do set -- $cfg
CRUSE_OVERLAP_BWD=$1 CRUSE_FWD_SIDE_SKIPS=$2 timeout 200 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bt.json 2>gpurun_out/bt.err
python -c "
import json;d=json.loads(open('gpurun_out/bt.json').read().strip().splitlines()[-1]);print('overlap_bwd=$1 fwd_side_skips=$2 MS',d['ms_per_step'],d['loss'],'e2e',d['e2e']['ms_per_step'])" >> gpurun_out/ms.txt 2>&1
done
cat gpurun_out/ms.txt
