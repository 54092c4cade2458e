mkdir -p gpurun_out
for bn in 0 192 160 128; do
NRL_ASTAT_BN=$bn timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_bn$bn.json 2> gpurun_out/bench_bn$bn.err
python - gpurun_out/bench_bn$bn.json $bn <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read())
    t={k[0]:k[1] for k in j["roofline"]["top_kernels_ms_per_step"]}
    print("astat BN",sys.argv[2], round(j["ms_per_step"],4), "gemm_ms", round(j["roofline"]["kernel_ms_per_step"],4), {k:v for k,v in t.items() if k in ("gemm in_proj","gemm out_proj dgrad","gemm in_proj dgrad")})
except Exception as e: print("ERR", e)
PY
done
