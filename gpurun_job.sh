mkdir -p gpurun_out
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -2 gpurun_out/bench_q.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_q.json").read())
print(j["ms_per_step"], j["value"], j["e2e"]["value"])
print([ (k["kernel"], k["ms_per_step"], round(k["frac_of_hbm_peak"],3)) for k in j["hbm_kernels"]])
PY
timeout 600 python experiments/module_profile.py naml 2>&1 | head -5
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_naml.py -m gpu -q --timeout 900 -x 2>&1 | tail -2
