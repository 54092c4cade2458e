mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:nrl_gemm -s 54 -c 18 -o gpurun_out/prof_gemm_r01f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_gemm.log 2>&1
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref.json
du -sh gpurun_out
