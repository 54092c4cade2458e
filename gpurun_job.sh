# end-of-round verification on ONE box: GPU tests, smoke, reference arm, full bench line, ncu launch list of one step
set -x
timeout 600 python -m pytest tests -m gpu -q --timeout 500 2>&1 | tail -8 > gpurun_out/r02_final2_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final2_smoke.log 2>&1
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r02_final2_bench_reference_arm.json 2> gpurun_out/r02_final2_ref.err
timeout 500 python bench.py > gpurun_out/r02_final2_bench.json 2> gpurun_out/r02_final2_bench.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final2_launches.csv python profiles/ncu_step.py > gpurun_out/r02_final2_ncu.log 2>&1
tail -3 gpurun_out/r02_final2_pytest_gpu.log; tail -2 gpurun_out/r02_final2_smoke.log; tail -c 600 gpurun_out/r02_final2_bench_reference_arm.json; echo; head -c 1500 gpurun_out/r02_final2_bench.json; echo; wc -l gpurun_out/r02_final2_launches.csv
