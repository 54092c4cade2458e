mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3
timeout 600 python experiments/module_profile.py > gpurun_out/module_profile.txt 2>&1; cat gpurun_out/module_profile.txt | head -12
timeout 600 python experiments/plm_profile.py 40 2>&1 | grep "wall\|pack_weights\|ce_bwd"
