mkdir -p gpurun_out/r2e
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline --no-train-block --steps 30 --table gpurun_out/r2e/kernels_infer.md | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('infer ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'loss', d['loss'])"; done
grep -E "stft|istft|wo_male" gpurun_out/r2e/kernels_infer.md | head -12
