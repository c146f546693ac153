timeout 900 python -m pytest tests/test_gpu_bwd.py -m gpu -q --timeout 600 -k "trainer or si_snr" 2>&1 | grep -E "^E |passed|failed|Error" | head -30
