mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
for v in 0 1; do
NRL_COMPACT_DGRAD=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_cd$v.json 2> gpurun_out/bench_cd$v.err
python - gpurun_out/bench_cd$v.json $v <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read())
    print("compact",sys.argv[2], round(j["ms_per_step"],4), round(j["value"]), "frac", round(j["roofline"]["frac"],4), [(k[0],k[1]) for k in j["roofline"]["top_kernels_ms_per_step"][:16]])
except Exception as e: print("ERR", e)
PY
done
