mkdir -p gpurun_out
N=${1:-8}
for ctas in 0 296 1184; do
  NRL_EXCHANGE_CTAS=$ctas timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
    bench.py --gpus $N --steps 30 --warmup 3 --no-extras > gpurun_out/bench_n${N}_c$ctas.json 2> gpurun_out/bench_n${N}_c$ctas.err
  python - gpurun_out/bench_n${N}_c$ctas.json $ctas <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    x=d.get("exchange") or {}
    print("ctas",sys.argv[2], round(d["ms_per_step"],4), round(d["value"]), x.get("ms_per_step"), x.get("timeline_us_last_step_rank0"))
except Exception as e:
    print("ERR", e)
PY
done
