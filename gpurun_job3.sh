mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ldtm experiments/ldtm_layout.cu && /tmp/ldtm | tee gpurun_out/ldtm_layout.txt
for d in 0 1 2 3; do
NRL_EPI_DEBUG=$d timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_dbg$d.json 2> gpurun_out/bench_dbg$d.err
done
python - <<'PY'
import json
for d in range(4):
    f=f"gpurun_out/bench_dbg{d}.json"
    try:
        j=json.loads(open(f).read())
        print(d, round(j["ms_per_step"],4), round(j["roofline"]["kernel_ms_per_step"],4), [(k[0],k[1]) for k in j["roofline"]["top_kernels_ms_per_step"] if k[0].startswith("gemm")])
    except Exception as e: print(f, "ERR", e)
PY
