mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
