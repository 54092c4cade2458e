mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tma.json 2> gpurun_out/bench_tma.err
NRL_ATTN_TMA=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_notma.json 2> gpurun_out/bench_notma.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_tma.json","gpurun_out/bench_notma.json"):
    try:
        d=json.loads(open(f).read())
        print(f, round(d["ms_per_step"],4), [(k[0],k[1]) for k in d["roofline"]["top_kernels_ms_per_step"][:8]])
    except Exception as e: print(f, "ERR", e)
PY
