mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -2 gpurun_out/bench_n4.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_n4.json").read().strip().splitlines()[-1])
print(j["n_gpus"], j["ms_per_step"], j["value"], j["impl_detail"]["exchange"][:40])
for k,v in j["configs"].items(): print(k, v.get("ms_per_step"), v.get("value"), v.get("error"))
PY
