mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s --timeout 900 2>&1 | tail -200 > gpurun_out/pytest_fullsize.log
tail -5 gpurun_out/pytest_fullsize.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --ignore tests/test_gpu_fullsize.py 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "from __graft_entry__ import smoke; smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:nrl_gemm_tc2 -f -o gpurun_out/r02a_gemm2 python profiles/ncu_step.py > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"attn_|pool_|gather_split|emb_grad|adam|dropout_words" -f -o gpurun_out/r02a_mem python profiles/ncu_step.py > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
