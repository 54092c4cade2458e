#!/usr/bin/env python
"""Dense GEMM peaks of the precisions the parity mode could use ("TF32 / 3xTF32 peak: not measured
yet -- builder must add it").  cuBLAS through torch.matmul, 8192^3, best of 10 (burst) and a 2-second back-to-back loop
(sustained), CUDA events.  bf16x3 / 3xTF32 are not library modes: they are three passes of the base precision, so their
ceilings are a third of the measured base rates (and bf16x3 additionally reads each operand tile once for three MMAs).

    gpurun -- 'python experiments/measure_peaks.py > gpurun_out/peaks.json'
"""
import json
import time

import torch

N = 8192
dev = torch.device("cuda", 0)


def bench(a, b, secs=2.0):
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    n, t0 = 0, time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(20):
            a @ b
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    flop = 2.0 * N ** 3
    return flop / (best * 1e-3) / 1e12, flop * n / (e0.elapsed_time(e1) * 1e-3) / 1e12


out = {"gpu": torch.cuda.get_device_name(0), "n": N}
a16 = torch.randn(N, N, device=dev, dtype=torch.bfloat16); b16 = torch.randn(N, N, device=dev, dtype=torch.bfloat16)
out["bf16_tflops"], out["bf16_tflops_sustained"] = bench(a16, b16)
a32 = torch.randn(N, N, device=dev); b32 = torch.randn(N, N, device=dev)
torch.backends.cuda.matmul.allow_tf32 = True
out["tf32_tflops"], out["tf32_tflops_sustained"] = bench(a32, b32)
torch.backends.cuda.matmul.allow_tf32 = False
out["fp32_simt_tflops"], out["fp32_simt_tflops_sustained"] = bench(a32, b32, secs=1.0)
out["bf16x3_ceiling_tflops_sustained"] = out["bf16_tflops_sustained"] / 3
out["tf32x3_ceiling_tflops_sustained"] = out["tf32_tflops_sustained"] / 3
print(json.dumps(out, indent=1))
