# end-of-round verification on ONE box at HEAD: GPU tests, smoke, reference arm, full bench line
set -x
timeout 600 python -m pytest tests -m gpu -q --timeout 500 2>&1 | tail -8 > gpurun_out/r02_final4_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final4_smoke.log 2>&1
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r02_final4_bench_reference_arm.json 2> gpurun_out/r02_final4_ref.err
timeout 500 python bench.py > gpurun_out/r02_final4_bench.json 2> gpurun_out/r02_final4_bench.err
tail -3 gpurun_out/r02_final4_pytest_gpu.log; tail -2 gpurun_out/r02_final4_smoke.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_final4_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["clocks"])
print({k:(round(v["ms_per_step"],3) if isinstance(v,dict) and "ms_per_step" in v else v) for k,v in d["configs"].items()})
print([k for k in d["roofline"]["top_kernels_ms_per_step"] if k[0]=="adam"])
r=json.loads(open("gpurun_out/r02_final4_bench_reference_arm.json").read().strip().splitlines()[-1]); print(r["value"], r["impl"], r["cpu_baseline"]["kind"])
PY
