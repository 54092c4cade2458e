mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 240 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
NRL_ATTN_SIMT=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err
cat gpurun_out/bench_simt.json | cut -c1-400
