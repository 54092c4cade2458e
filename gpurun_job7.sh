mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err
tail -5 gpurun_out/bench_r02c.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r02c.json").read())
print(round(j["ms_per_step"],4), round(j["value"]), "e2e", round(j["e2e"]["value"]), "launches", j["gpu_launches"], "frac", round(j["roofline"]["frac"],4))
print(j["cpu_baseline"])
for k,v in (j.get("configs") or {}).items(): print(k, {kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk!="workload"})
PY
