mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python profiles/ncu_step.py > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:nrl_gemm_tc -f -o gpurun_out/r02_gemm python profiles/ncu_step.py > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"attn_|pool_|gather_split|emb_grad|adam|dropout_words|score_loss|pack_weights" -f -o gpurun_out/r02_mem python profiles/ncu_step.py > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err; cut -c1-200 gpurun_out/bench_r02e.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_r02.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_r02.json
