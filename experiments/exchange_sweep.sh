# Sweep the grid of the fused gradient-exchange kernel and compare with the NCCL path.
# Usage on a multi-GPU box:   gpurun --gpus N -- 'bash experiments/exchange_sweep.sh N'
# Each run prints: exchange, CTAs, ms/step, impressions/s, exchange-kernel ms, GB/s per NVLink direction.
N=${1:-2}
mkdir -p gpurun_out
run() {  # $1 = exchange, $2 = NRL_EXCHANGE_CTAS
  NRL_EXCHANGE_CTAS=$2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29621 bench.py --gpus $N --steps 40 --warmup 3 --exchange $1 > gpurun_out/sweep_$1_$2.json 2> gpurun_out/sweep_$1_$2.err
  python - "$1" "$2" gpurun_out/sweep_$1_$2.json <<'PY'
import json, sys
ex, ctas, path = sys.argv[1:4]
try:
    d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    x = d.get("exchange") or {}
    print(f"{ex:5s} ctas={ctas:>5s}  {d['ms_per_step']:.3f} ms/step  {d['value']:,.0f} imp/s  "
          f"exchange kernel {x.get('ms_per_step')} ms  {x.get('achieved_gbs_per_direction')} GB/s/dir")
except Exception as e:
    print(f"{ex:5s} ctas={ctas:>5s}  FAILED: {e}")
PY
}
run nccl 0
for c in 148 296 592 1184 2368; do run peer $c; done
