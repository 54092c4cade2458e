"""Mint ``tests/golden/metrics_ref.npz``: the reference's OWN aspect-based diversity / personalization functions
(``newsreclib/metrics/functional.py:8-49,52-110``, imported unmodified from ``/root/reference``) on seeded impressions.

Runs ONLY in the build container.  ``functional.py`` imports one helper from torchmetrics (absent here, third-party):
``torchmetrics.utilities.checks._check_retrieval_functional_inputs``, an input check that flattens its two arguments and
casts them (preds -> float, target -> long with ``allow_non_binary_target=True``); the stand-in below does exactly that.
The per-impression grouping and the mean over impressions are torchmetrics' ``RetrievalMetric.compute`` /
``newsreclib/metrics/base.py:144-181`` (group by ``indexes``; a group whose aspect ids sum to 0 counts 0.0 --
``empty_target_action="neg"``; mean over the groups), restated in ``grouped`` below -- "parity unpinned" for that part.

Usage:  python oracle/make_metrics_golden.py [--check]
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle._golden_io import save as golden_save  # noqa: E402


def _check_retrieval_functional_inputs(preds, target, allow_non_binary_target=False):
    return preds.flatten().float(), target.flatten().long()


for name in ("torchmetrics", "torchmetrics.utilities", "torchmetrics.utilities.checks"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["torchmetrics.utilities.checks"]._check_retrieval_functional_inputs = _check_retrieval_functional_inputs

from newsreclib.metrics.functional import diversity, personalization  # noqa: E402  (the reference's own file)


def grouped(fn, sizes, *per_candidate, hist=None, hist_sizes=None):
    res, o, ho = [], 0, 0
    for i, c in enumerate(sizes):
        args = [t[o:o + c] for t in per_candidate]
        if hist is not None:
            args.append(hist[ho:ho + hist_sizes[i]])
            ho += hist_sizes[i]
        res.append(float(fn(*args)) if int(args[1].sum()) != 0 else 0.0)
        o += c
    return np.array(res, dtype=np.float64)


if __name__ == "__main__":
    rng = np.random.default_rng(7)
    n_imp, ncat, nsent = 37, 19, 4          # num_categ_classes + 1, num_sent_classes + 1 (nrms_module.py:110-111)
    sizes = rng.integers(1, 40, n_imp)
    hist_sizes = rng.integers(1, 50, n_imp)
    N, NH = int(sizes.sum()), int(hist_sizes.sum())
    preds = torch.from_numpy(rng.permutation(N).astype(np.float32) / N)           # distinct scores: no tie ambiguity
    cat, sent = torch.from_numpy(rng.integers(1, ncat, N)), torch.from_numpy(rng.integers(1, nsent, N))
    hcat, hsent = torch.from_numpy(rng.integers(1, ncat, NH)), torch.from_numpy(rng.integers(1, nsent, NH))
    cat[:sizes[0]] = 0                                                             # one impression with "no target"
    rec = dict(sizes=sizes, hist_sizes=hist_sizes, preds=preds.numpy(), cat=cat.numpy(), sent=sent.numpy(),
               hcat=hcat.numpy(), hsent=hsent.numpy(), num_classes=np.array([ncat, nsent]))
    for k in (5, 10):
        rec[f"categ_div@{k}"] = grouped(lambda p, t: diversity(p, t, ncat, top_k=k), sizes, preds, cat)
        rec[f"sent_div@{k}"] = grouped(lambda p, t: diversity(p, t, nsent, top_k=k), sizes, preds, sent)
        rec[f"categ_pers@{k}"] = grouped(lambda p, t, h: personalization(p, t, h, ncat, top_k=k), sizes, preds, cat,
                                         hist=hcat, hist_sizes=hist_sizes)
        rec[f"sent_pers@{k}"] = grouped(lambda p, t, h: personalization(p, t, h, nsent, top_k=k), sizes, preds, sent,
                                        hist=hsent, hist_sizes=hist_sizes)
    path = os.path.join(ROOT, "tests", "golden", "metrics_ref.npz")
    golden_save(path, **rec)
    print(f"[metrics_ref] {n_imp} impressions, {N} candidates; categ_div@5 mean {rec['categ_div@5'].mean():.4f}, "
          f"categ_pers@10 mean {rec['categ_pers@10'].mean():.4f}; {os.path.getsize(path)} bytes")
