mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tfm.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_naml.py -m gpu -q --timeout 600 2>&1 | tail -4
timeout 600 python experiments/plm_profile.py 96 > gpurun_out/plm_profile_96.txt 2>&1; cat gpurun_out/plm_profile_96.txt | head -22
NRL_GEMM_FINE=0 timeout 600 python experiments/plm_profile.py 96 2>&1 | head -3
timeout 600 python experiments/plm_profile.py 40 > gpurun_out/plm_profile_40.txt 2>&1; cat gpurun_out/plm_profile_40.txt | head -3
python __graft_entry__.py --smoke 2>&1 | tail -3
