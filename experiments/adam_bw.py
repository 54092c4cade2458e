"""Adam pass (nrl_adam_step_zero_grad) over flat buffers of the NRMS (21.8 M) and NAML (40 M elements) sizes for
different grid sizes (NRL_ADAM_CTAS_PER_SM): effective GB/s over the 8 streams (p, g, m, v read; p, m, v, g written).
Buffers larger than L2 are rotated so that no pass starts on a warm L2.  Usage: python experiments/adam_bw.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from newsreclib_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    for n in (21_843_500, 40_000_000):
        sets = [[torch.randn(n, device=dev) for _ in range(4)] for _ in range(3)]
        for v in sets:
            v[3].abs_()
        for per_sm in ("default", "2", "3", "4", "5", "6", "8", "12", "16"):
            if per_sm == "default":
                os.environ.pop("NRL_ADAM_CTAS_PER_SM", None)
            else:
                os.environ["NRL_ADAM_CTAS_PER_SM"] = per_sm
            for i in range(3):
                ops.adam_step(*sets[i % 3], 1 + i, zero_grad=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            reps = 30
            for i in range(reps):
                ops.adam_step(*sets[i % 3], 4 + i, zero_grad=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(f"n = {n / 1e6:5.1f} M  CTAs/SM {per_sm:>7}: {ms * 1e3:7.1f} us  {8 * 4 * n / ms / 1e6:7.0f} GB/s")


if __name__ == "__main__":
    main()
