mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4 > gpurun_out/r02j_pytest_gpu.log; tail -3 gpurun_out/r02j_pytest_gpu.log
timeout 600 python experiments/plm_profile.py 40 2>&1 | grep "wall\|pack_weights"
timeout 1200 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err; tail -2 gpurun_out/bench_r02j.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r02j.json").read())
print(j["ms_per_step"], j["value"], j["e2e"]["value"], j["roofline"]["frac"], j["cpu_baseline"]["value"])
for k,v in j["configs"].items(): print(k, v.get("ms_per_step"), v.get("value"), v.get("error"))
p=j["configs"]["nrms_plm_roberta_base"]
print(json.dumps(p.get("roofline")))
print({k:(v.get("ms_per_step") if isinstance(v,dict) and "ms_per_step" in v else {kk:vv.get("ms_per_step") for kk,vv in v.items()} if isinstance(v,dict) else None) for k,v in p.items() if isinstance(v,dict) and k!="roofline"})
PY
