mkdir -p gpurun_out
for mb in 0 32 8; do
NRL_EXCHANGE_CHUNK_MB=$mb timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$mb bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2_$mb.json 2> gpurun_out/bench_n2.err
python -c "
import json
txt=open('gpurun_out/bench_n2_$mb.json').read().splitlines()
print(len(txt),'stdout lines')
d=json.loads([l for l in txt if l.startswith('{')][0]); print($mb, d['n_gpus'], round(d['ms_per_step'],3), round(d['value']))"
done
