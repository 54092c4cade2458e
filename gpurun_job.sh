timeout 600 python -m pytest tests/test_gpu_tfm.py -m gpu -q --timeout 500 -k "integration_md" 2>&1 | tail -15
