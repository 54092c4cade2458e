mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tfm.py tests/test_gpu_naml.py -m gpu -q --timeout 600 -s -k "plm or tfm" 2>&1 | grep -v "^$" | tail -70 > gpurun_out/r02i_pytest_plm.log; grep "tfm\]\|plm head\]\|passed\|failed\|Error\|error" gpurun_out/r02i_pytest_plm.log | cut -c1-220 | head -40
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'tfm|nrl_gemm' -o gpurun_out/tfm python profiles/ncu_tfm.py > gpurun_out/ncu_tfm.log 2>&1; tail -3 gpurun_out/ncu_tfm.log
ncu -i gpurun_out/tfm.ncu-rep --page raw --csv > gpurun_out/r02_tfm_raw.csv 2>/dev/null; wc -l gpurun_out/r02_tfm_raw.csv; ls -la gpurun_out/tfm.ncu-rep
