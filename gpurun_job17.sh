mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.log
python -c "from __graft_entry__ import smoke; smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; cut -c1-160 gpurun_out/r02_bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python profiles/ncu_step.py > /dev/null 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:nrl_gemm_tc -f -o /tmp/r02_gemm python profiles/ncu_step.py > /dev/null 2>&1
ncu -i /tmp/r02_gemm.ncu-rep --page raw --csv > gpurun_out/r02_gemm_raw.csv 2>/dev/null
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"attn_|pool_|gather_split|emb_grad|adam|dropout_words|score_loss|pack_weights|dense_" -f -o /tmp/r02_mem python profiles/ncu_step.py > /dev/null 2>&1
ncu -i /tmp/r02_mem.ncu-rep --page raw --csv > gpurun_out/r02_mem_raw.csv 2>/dev/null
ls -la gpurun_out/ | head -30; du -sh gpurun_out
