mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
for v in 0 1; do
NRL_WGRAD_STREAM=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_ws$v.json 2> gpurun_out/bench_ws$v.err
python - gpurun_out/bench_ws$v.json $v <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read())
    print("wgrad_stream",sys.argv[2], round(j["ms_per_step"],4), round(j["value"]), "e2e", round(j["e2e"]["ms_per_step"],4), "eval", round(j["eval_forward"]["ms_per_step"],4))
except Exception as e: print("ERR", e)
PY
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r02d.json").read())
for k,v in (j.get("configs") or {}).items(): print(k, {kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk!="workload"})
PY
