#!/usr/bin/env python
"""The PLM head (dropout -> MHSA across the N news of the call -> dropout -> additive pooling, text.py:93-100) at the
shape of one NRMS-PLM training step's history call (N = 400 news, T = 28 tokens, 768-d, 16 heads of 48), forward +
backward, between cudaProfilerStart/Stop for ncu (flash-style attention kernels of nrl_attn_flash.cuh, TMA-staged pooling):

    ncu --profile-from-start off --set full --clock-control none -k regex:'attn|pool' -o /tmp/head python profiles/ncu_plm_head.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from newsreclib_b200 import ops  # noqa: E402

N, T, E, H, Q = 400, 28, 768, 16, 200
g = torch.Generator().manual_seed(0)
shapes = [(3 * E, E), (3 * E,), (E, E), (E,), (Q, E), (Q,), (Q,)]
params = [(torch.randn(*s, generator=g) * (0.04 if len(s) == 2 else 0.1)).cuda().requires_grad_(True) for s in shapes]
x = torch.randn(N, T, E, generator=g).cuda().requires_grad_(True)
w = torch.randn(N, E, generator=g).cuda()


def step(seed):
    out = ops.PlmHeadFn.apply(x, *params, H, 0, 0.2, True, seed, ops.PREC_BF16X3)
    (out * w).sum().backward()


step(1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one PLM head forward + backward")
