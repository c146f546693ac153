B() { env $2 $3 $4 timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'loss', d['loss'])"; }
B default X=1
B head4 CRUSE_HEAD_CHUNKS=4
B head5 CRUSE_HEAD_CHUNKS=5
B head6 CRUSE_HEAD_CHUNKS=6
B head4cap60 CRUSE_HEAD_CHUNKS=4 CRUSE_HEAD_CAP=60
B head5cap60 CRUSE_HEAD_CHUNKS=5 CRUSE_HEAD_CAP=60
B head3cap84 CRUSE_HEAD_CHUNKS=3
