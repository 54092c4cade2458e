mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tfm.py -m gpu -q --timeout 600 -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02f_pytest_tfm.log; grep "tfm\]\|passed\|failed" gpurun_out/r02f_pytest_tfm.log
timeout 1200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f3.json 2> gpurun_out/bench_f3.err; tail -3 gpurun_out/bench_f3.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_f3.json").read())
print(j["ms_per_step"], j["value"])
print(json.dumps(j["configs"].get("nrms_plm_roberta_base"), indent=1))
PY
