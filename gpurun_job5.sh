mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/p2p_bw experiments/p2p_bw.cu && /tmp/p2p_bw | tee gpurun_out/p2p_bw.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5
NRL_ATTN_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -2
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r02b.json").read())
print(round(j["ms_per_step"],4), j["gpu_launches"], round(j["roofline"]["kernel_ms_per_step"],4), [(k[0],k[1]) for k in j["roofline"]["top_kernels_ms_per_step"]])
PY
bash gpurun_job_n2.sh auto nccl
