mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_naml.py tests/test_gpu_tfm.py -m gpu -q --timeout 600 -s -k "plm or tfm" 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02g_pytest_plm.log; grep "tfm\]\|plm head\]\|passed\|failed\|Error\|error" gpurun_out/r02g_pytest_plm.log | head -40
timeout 600 python experiments/plm_profile.py 40 > gpurun_out/plm_profile_40.txt 2>&1; cat gpurun_out/plm_profile_40.txt | head -16
timeout 600 python experiments/plm_profile.py 96 > gpurun_out/plm_profile_96.txt 2>&1; cat gpurun_out/plm_profile_96.txt | head -16
