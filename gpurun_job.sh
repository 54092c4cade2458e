mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'pool|attn|gather|emb_grad|adam|dropout|dense|score|pack' -o /tmp/mem python profiles/ncu_step.py > gpurun_out/ncu_mem.log 2>&1; tail -2 gpurun_out/ncu_mem.log
ncu -i /tmp/mem.ncu-rep --page raw --csv > gpurun_out/r02k_mem_raw.csv 2>/dev/null; wc -l gpurun_out/r02k_mem_raw.csv
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'attn|pool' -o /tmp/head python profiles/ncu_plm_head.py > gpurun_out/ncu_head.log 2>&1; tail -2 gpurun_out/ncu_head.log
ncu -i /tmp/head.ncu-rep --page raw --csv > gpurun_out/r02k_head_raw.csv 2>/dev/null; wc -l gpurun_out/r02k_head_raw.csv
