# A/B of the training-step schedule on one GPU: eager vs CUDA graph, weight gradients serial vs beside the BPTT
: > gpurun_out/ms.txt
for cfg in "0 1 0 --no-graph" "0 1 0" "1 1 0" "1 1 1" "1 0 0"; do set -- $cfg
CRUSE_OVERLAP_BWD=$1 CRUSE_BWD_SIDE_CAP=$2 CRUSE_BWD_SIDE_L1=$3 timeout 200 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline $4 > gpurun_out/bt_$1$2$3$4.json 2>gpurun_out/bt_$1$2$3$4.err
python -c "
import json;d=json.loads(open('gpurun_out/bt_$1$2$3$4.json').read().strip().splitlines()[-1]);print('overlap=$1 cap=$2 l1side=$3 $4 MS',d['ms_per_step'],d['loss'],'e2e',d['e2e']['ms_per_step'])" >> gpurun_out/ms.txt 2>&1
done
cat gpurun_out/ms.txt
timeout 400 python -m pytest tests/test_gpu_bwd.py -m gpu -x -q > gpurun_out/pt.log 2>&1; tail -n 3 gpurun_out/pt.log
