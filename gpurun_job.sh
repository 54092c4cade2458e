mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3 > gpurun_out/r02_final_pytest_gpu.log; cat gpurun_out/r02_final_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/r02_final_smoke.log 2>&1; tail -2 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/r02_final_bench_reference_arm.json 2> gpurun_out/ref.err; tail -c 300 gpurun_out/r02_final_bench_reference_arm.json
timeout 1200 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; tail -2 gpurun_out/r02_final_bench.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02_final_bench.json").read())
print(j["ms_per_step"], j["value"], j["e2e"]["value"], j["roofline"]["frac"], j["cpu_baseline"]["value"], j["gpu_launches"], j["clocks"])
for k in j["roofline"]["per_gemm"]: print("  ", k)
for k in j["hbm_kernels"]: print("  ", k["kernel"], k["ms_per_step"], k["frac_of_hbm_peak"])
for k,v in j["configs"].items(): print(k, v.get("ms_per_step"), v.get("value"), v.get("error"))
PY
