mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -3
for v in 0 1; do
NRL_ATTN_BWD64=$v timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_u$v.json 2> gpurun_out/bench_u$v.err; tail -1 gpurun_out/bench_u$v.err
python - $v <<'PY'
import json,sys
j=json.loads(open(f"gpurun_out/bench_u{sys.argv[1]}.json").read())
print("ldsm64", sys.argv[1], j["ms_per_step"], j["value"])
for k in j["hbm_kernels"]:
    if "attn_bwd" in k["kernel"]: print("   ", k["kernel"], k["ms_per_step"], k["frac_of_hbm_peak"])
PY
done
