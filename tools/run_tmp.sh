timeout 600 python -m pytest tests/test_gpu_frontend.py -m gpu -q --timeout 600 2>&1 | grep -E "^E |passed|failed|Error" | head -20
