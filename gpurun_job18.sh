mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_modules.py tests/test_gpu_reference_goldens.py -m gpu -q --timeout 600 2>&1 | tail -3
for v in 0 1; do
NRL_QKV_PLANES=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_qp$v.json 2> gpurun_out/bench_qp$v.err
python - gpurun_out/bench_qp$v.json $v <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read())
    t={k[0]:k[1] for k in j["roofline"]["top_kernels_ms_per_step"]}
    print("qkv planes",sys.argv[2], round(j["ms_per_step"],4), round(j["value"]), "eval", round(j["eval_forward"]["ms_per_step"],4), {k:v for k,v in t.items() if k in ("gemm in_proj","attn_fwd","attn_bwd")})
except Exception as e: print("ERR", e)
PY
tail -2 gpurun_out/bench_qp$v.err
done
