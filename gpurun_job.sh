mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tfm.py -m gpu -q --timeout 600 -s 2>&1 | grep "tfm\]\|passed\|failed\|Error" | cut -c1-260
