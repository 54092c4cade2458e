mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print(j["n_gpus"], j["ms_per_step"], j["value"], j["impl_detail"], j.get("exchange"))
for k,v in j["configs"].items(): print(k, v.get("ms_per_step"), v.get("value"), v.get("error"))
print(json.dumps(j["configs"]["nrms_plm_roberta_base"].get("roofline")))
PY
timeout 600 python -m pytest tests/test_gpu_peer_exchange.py -m gpu -q --timeout 500 2>&1 | tail -3
