mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tfm.py tests/test_gpu_naml.py -m gpu -q --timeout 600 -s -k "module" 2>&1 | grep "tfm\]\|passed\|failed\|Error" | cut -c1-300
timeout 600 python experiments/plm_profile.py 40 > gpurun_out/plm_profile_40.txt 2>&1; cat gpurun_out/plm_profile_40.txt | head -24
timeout 600 python experiments/plm_profile.py 96 > gpurun_out/plm_profile_96.txt 2>&1; cat gpurun_out/plm_profile_96.txt | head -4
