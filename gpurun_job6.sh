mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer_exchange.py -m gpu -q -x --timeout 400 2>&1 | tail -8
for ex in auto nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 40 --warmup 3 --exchange $ex > gpurun_out/bench_n2_$ex.json 2> gpurun_out/bench_n2_$ex.err
  echo "== $ex rc=$?"; python - gpurun_out/bench_n2_$ex.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(round(d["ms_per_step"],4), round(d["value"]), d["impl_detail"]["exchange"][:60], d.get("exchange"))
except Exception as e:
    print("ERR", e)
PY
  tail -3 gpurun_out/bench_n2_$ex.err
done
