timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k fused --timeout 600 2>&1 | grep -E "assert|Error|error|passed|failed|^E " | head -30
