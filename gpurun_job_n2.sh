# two-GPU job: the exchange tests (incl. the two-process CUDA-IPC run) and the 2-GPU bench (default --exchange auto;
# add `nccl` / `peer` as arguments to A/B the exchanges)   (run with: gpurun --gpus 2 -- 'bash gpurun_job_n2.sh [modes]')
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_peer_exchange.py -m gpu -q -x --timeout 300 2>&1 | tail -15 > gpurun_out/pytest_exchange.log; cat gpurun_out/pytest_exchange.log
for ex in ${@:-auto}; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 40 --warmup 3 --exchange $ex > gpurun_out/bench_n2_$ex.json 2> gpurun_out/bench_n2_$ex.err
  echo "== $ex rc=$?"; cut -c1-300 gpurun_out/bench_n2_$ex.json; tail -3 gpurun_out/bench_n2_$ex.err
done
