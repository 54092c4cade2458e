mkdir -p gpurun_out
NRL_EPI_DEBUG=8 timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -4
for d in 0 8; do
NRL_EPI_DEBUG=$d timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_rel$d.json 2> gpurun_out/bench_rel$d.err
done
python - <<'PY'
import json
for d in (0,8):
    f=f"gpurun_out/bench_rel{d}.json"
    try:
        j=json.loads(open(f).read())
        print(d, round(j["ms_per_step"],4), round(j["roofline"]["kernel_ms_per_step"],4), [(k[0],k[1]) for k in j["roofline"]["top_kernels_ms_per_step"]])
    except Exception as e: print(f, "ERR", e)
PY
