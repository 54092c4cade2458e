#!/usr/bin/env python
"""One roberta-base-shaped encoder layer (forward + backward, trainable, all dropouts on) of the sm_100a PLM transformer
between cudaProfilerStart/Stop, for ncu:

    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/tfm \
        python profiles/ncu_tfm.py [N] [T]

Default: 400 titles x 48 tokens = 19 200 token rows (the row-streaming GEMMs run on the CTA-pair kernel).  One untimed
warm-up pass first.  Numbers printed under ncu are never bench values."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from newsreclib_b200 import ops  # noqa: E402
from tfm_helpers import param_list, random_text, random_tfm_params  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
T = int(sys.argv[2]) if len(sys.argv) > 2 else 48
cfg = dict(hidden=768, heads=12, inter=3072, layers=1, vocab=5000, max_pos=T + 4, eps=1e-5)
P = random_tfm_params(768, 12, 3072, 1, 5000, T + 4, seed=1, wstd=0.03)
ids, att = random_text(N, T, 5000, seed=2, min_len=T // 3)
st = ops.TfmState(768, 12, 3072, 1, 5000, T + 4, 1, 1e-5, 0.1, 0.1)
leaves, _ = param_list(P, 1, "cuda")
ids, att = ids.cuda(), att.cuda()
w = torch.randn(N, T, 768, device="cuda")


def step(seed):
    for t in leaves:
        t.grad = None
    out = ops.TfmEncoderFn.apply(ids, att, st, True, seed, ops.PREC_BF16X3, *leaves)
    (out * w).sum().backward()


step(1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"profiled one layer, {N * T} token rows")
