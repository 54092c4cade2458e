"""Per-impression ranking metrics on device (AUC, MRR, nDCG@k) for the epoch-end hooks.

The reference computes them with torchmetrics (``nrms_module.py:182-191``: ``AUROC(task="binary")`` over all
candidates of the epoch, ``RetrievalMRR()`` and ``RetrievalNormalizedDCG(top_k=k)`` with ``indexes`` = the
impression of each candidate), which is not installed here; these are vectorised restatements of those
definitions (third-party, torchmetrics 1.x):

* retrieval metrics are computed per impression and averaged over ALL impressions; an impression without a
  positive counts as 0 (torchmetrics' default ``empty_target_action="neg"``);
* nDCG uses the labels as gains, discount ``1 / log2(rank + 1)``, and the ideal ordering of the same impression;
* AUROC is the global binary AUROC over every candidate, ties in the scores counted half (mid-ranks), which is
  what the trapezoidal ROC integral gives.

One dense ``[impressions, max candidates]`` sort, no python loop over impressions, no host sync except the
width; not part of the timed hot path.
"""
from __future__ import annotations

from typing import Dict, List

import torch


def _dense(preds, targets, sizes):
    B, M = sizes.numel(), int(sizes.max())
    idx = torch.arange(M, device=preds.device)[None, :] < sizes[:, None]
    p = torch.full((B, M), float("-inf"), device=preds.device)
    t = torch.zeros(B, M, device=preds.device)
    p[idx] = preds.float()
    t[idx] = targets.float()
    return p, t, idx


def binary_auroc(preds: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """Mann-Whitney form of the binary AUROC with mid-ranks for tied scores (float64 rank sums: an epoch of
    MINDlarge-dev is 1.4e7 candidates)."""
    pf, pos = preds.float().reshape(-1), targets.reshape(-1) > 0
    _, inv, counts = torch.unique(pf, sorted=True, return_inverse=True, return_counts=True)
    counts = counts.double()
    mid = counts.cumsum(0) - (counts - 1.0) / 2.0          # 1-based average rank of each distinct score
    r = mid[inv]
    npos = pos.sum().double()
    nneg = pos.numel() - npos
    auc = (r[pos].sum() - npos * (npos + 1.0) / 2.0) / (npos * nneg).clamp_min(1.0)
    return auc.float()


def ranking_metrics(preds: torch.Tensor, targets: torch.Tensor, sizes: torch.Tensor,
                    top_k_list: List[int]) -> Dict[str, torch.Tensor]:
    """preds/targets: concatenated per-impression scores/labels; sizes: candidates per impression."""
    sizes = sizes.to(preds.device)
    p, t, _ = _dense(preds, targets, sizes)
    order = p.argsort(dim=1, descending=True, stable=True)
    ts = t.gather(1, order)
    ranks = torch.arange(1, p.shape[1] + 1, device=p.device, dtype=torch.float32)[None, :]
    has_pos = ts.sum(1) > 0
    first = (ts > 0).float().argmax(dim=1).float() + 1
    out = {"mrr": torch.where(has_pos, 1.0 / first, torch.zeros_like(first)).mean()}
    disc = 1.0 / torch.log2(ranks + 1)
    ideal = t.sort(dim=1, descending=True).values
    for k in top_k_list:
        dcg = (ts[:, :k] * disc[:, :k]).sum(1)
        idcg = (ideal[:, :k] * disc[:, :k]).sum(1)
        out[f"ndcg@{k}"] = torch.where(idcg > 0, dcg / idcg.clamp_min(1e-12), torch.zeros_like(dcg)).mean()
    out["auc"] = binary_auroc(preds, targets)
    return out
