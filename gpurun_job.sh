mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15 > gpurun_out/r02h_pytest_gpu.log; tail -5 gpurun_out/r02h_pytest_gpu.log
timeout 600 python experiments/plm_profile.py 40 > gpurun_out/plm_profile_40.txt 2>&1; cat gpurun_out/plm_profile_40.txt | head -12
timeout 1200 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err; tail -3 gpurun_out/bench_r02h.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r02h.json").read())
print(j["ms_per_step"], j["value"], j["e2e"]["value"], j["roofline"]["frac"], j["cpu_baseline"]["value"])
for k,v in j["configs"].items(): print(k, v.get("ms_per_step"), v.get("value"))
print(json.dumps(j["configs"].get("nrms_plm_roberta_base"))[:1500])
PY
