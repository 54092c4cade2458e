mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 2>&1 | tail -3
for v in 0 1 2; do
NRL_ATTN_BWD_VARIANT=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_ab$v.json 2> gpurun_out/bench_ab$v.err
python - gpurun_out/bench_ab$v.json $v <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read())
    t={k[0]:k[1] for k in j["roofline"]["top_kernels_ms_per_step"]}
    print("attn_bwd variant",sys.argv[2], round(j["ms_per_step"],4), round(j["value"]), "attn_bwd", t.get("attn_bwd"), "attn_fwd", t.get("attn_fwd"))
except Exception as e: print("ERR", e)
PY
done
NRL_ATTN_BWD_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -2
