mkdir -p gpurun_out
N=${1:-8}
run() {  # $1 = tag, rest = bench args
  tag=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
    bench.py --gpus $N --steps 30 --warmup 3 "$@" > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
  echo "== $tag rc=$?"
  python - gpurun_out/bench_n${N}_$tag.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["value"]), d["impl_detail"]["exchange"][:50], d.get("exchange"))
    for k,v in (d.get("configs") or {}).items(): print("   ", k, {kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk!="workload"})
except Exception as e:
    print("ERR", e)
PY
  tail -2 gpurun_out/bench_n${N}_$tag.err
}
run auto
run nccl --exchange nccl --no-extras
