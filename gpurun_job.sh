mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -2
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_x.json; python -c "import json; d=json.load(open('gpurun_out/bench_x.json')); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_step'],3), [x for x in d['roofline']['top_kernels_ms_per_step'] if not x[0].startswith('gemm')])"
