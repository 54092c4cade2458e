mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -3
for v in 0 1; do
NRL_POOL_TMA=$v timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_p$v.json 2> gpurun_out/bench_p$v.err; tail -1 gpurun_out/bench_p$v.err
python - $v <<'PY'
import json,sys
j=json.loads(open(f"gpurun_out/bench_p{sys.argv[1]}.json").read())
print("pool tma", sys.argv[1], j["ms_per_step"], j["value"], j["eval_forward"]["ms_per_step"])
for k in j["hbm_kernels"]: print("   ", k["kernel"], k["ms_per_step"], k["frac_of_hbm_peak"])
PY
done
