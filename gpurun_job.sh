timeout 300 python -m pytest tests/test_gpu_naml.py -m gpu -q --timeout 240 -k module_trainer 2>&1 | grep -E "Error|assert|error" | head -12
