"""GPU time of the SMALL NT GEMMs of the NRMS user block (3 200 rows: in-projection N = 900, out-projection N = 300,
additive N = 200; K = 304) for different n-tile caps (NRL_GEMM_BN_MAX, read once per process -> one subprocess per value):
these launches are latency-bound (one tile per CTA, K = 5 k-blocks through a 2-stage ring at BN = 256).
Usage: python experiments/small_gemm_latency.py"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    import torch
    sys.path.insert(0, ROOT)
    from newsreclib_b200 import _lib, ops
    lib = _lib.load()
    dev = torch.device("cuda:0")
    for M, N, K in ((3200, 900, 304), (3200, 300, 304), (3200, 200, 304), (3200, 304, 208), (105600, 900, 304)):
        A, B = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
        for _ in range(5):
            ops.gemm_test(A, B, False)
        torch.cuda.synchronize()
        lib.nrl_profile_start(torch.cuda.current_stream().cuda_stream)
        reps = 40
        for _ in range(reps):
            ops.gemm_test(A, B, False)
        torch.cuda.synchronize()
        names = C.create_string_buffer(48 * 4 * reps + 48)
        ms = (C.c_float * (4 * reps + 1))()
        n = lib.nrl_profile_stop(names, 48, ms, 4 * reps)
        ts = [ms[i] for i in range(n) if names.raw[i * 48:(i + 1) * 48].split(b"\0")[0].decode().startswith("gemm_test")]
        ts.sort()
        print(f"  BN_MAX {os.environ.get('NRL_GEMM_BN_MAX', '256'):>4}  M {M:6d} N {N:4d} K {K}: median {1e3 * ts[len(ts) // 2]:7.1f} us  min {1e3 * ts[0]:7.1f} us")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for bn in ("256", "128", "96", "64"):
            env = dict(os.environ, NRL_GEMM_BN_MAX=bn)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=False)
