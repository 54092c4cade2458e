mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_tfm.py -m gpu -q --timeout 500 2>&1 | tail -2
timeout 600 python experiments/plm_profile.py 40 2>&1 | grep "wall\|ln._bwd"
timeout 600 python experiments/plm_profile.py 96 2>&1 | grep "wall\|ln._bwd"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -q --timeout 1400 --deselect tests/test_gpu_peer_exchange.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep "passed\|failed\|ERROR SUMMARY" gpurun_out/r02_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_tfm.py tests/test_gpu_naml.py tests/test_gpu_parity.py -m gpu -q --timeout 1100 > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep "passed\|failed\|RACECHECK SUMMARY" gpurun_out/r02_sanitizer_racecheck.log
