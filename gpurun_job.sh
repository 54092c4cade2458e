mkdir -p gpurun_out
NRL_ATTN_BWD_VARIANT=3 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_reference_goldens.py tests/test_gpu_modules.py -m gpu -q --timeout 900 2>&1 | tail -6
for v in 1 3; do
NRL_ATTN_BWD_VARIANT=$v timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; tail -1 gpurun_out/bench_v$v.err
python - $v <<'PY'
import json,sys
j=json.loads(open(f"gpurun_out/bench_v{sys.argv[1]}.json").read())
t={k[0]:k[1] for k in j["roofline"]["top_kernels_ms_per_step"]}
print("variant", sys.argv[1], j["ms_per_step"], j["value"], "attn_bwd", t.get("attn_bwd"))
PY
done
