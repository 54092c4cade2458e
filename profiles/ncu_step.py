#!/usr/bin/env python
"""One NRMS training step of the bench.py workload between cudaProfilerStart/Stop, for ncu:

    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/step \
        python profiles/ncu_step.py [--precision bf16x3|bf16] [--eval]

Three untimed warm-up steps run first (tensor maps, attribute setup, allocator), then ONE step is profiled.
Numbers printed under ncu are never bench values.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from newsreclib_b200 import ops  # noqa: E402
from newsreclib_b200.synthetic import make_batch, make_nrms_params  # noqa: E402
from newsreclib_b200.trainer import NRMSTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--eval", action="store_true")
ap.add_argument("--vocab", type=int, default=70000)
ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
prec = ops.PREC_BF16X3 if a.precision == "bf16x3" else ops.PREC_BF16
tr = NRMSTrainer(make_nrms_params(a.vocab, seed=1234), 15, device=dev, dropout_p=0.2, precision=prec, status_every=0)
bs = []
for i in range(4):
    hb = make_batch(a.batch, a.vocab, hist="fixed", cand="train", seed=1234 + i)
    bs.append({"x_hist": {"title": hb["x_hist"]["title"].to(dev)}, "x_cand": {"title": hb["x_cand"]["title"].to(dev)},
               "batch_hist": hb["batch_hist"].to(dev), "batch_cand": hb["batch_cand"].to(dev), "labels": hb["labels"].to(dev)})
step = (lambda i: tr.eval_forward(bs[i % 4], a.batch, 50, 5)) if a.eval else (lambda i: tr.train_step(bs[i % 4], a.batch, 50, 5))
for i in range(3):
    step(i)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(3)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step")
