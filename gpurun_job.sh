# the standard GPU job of this repo: parity tests, smoke, headline bench (run with: gpurun -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "from __graft_entry__ import smoke; smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.json
