"""Per-impression ranking metrics on device (AUC, MRR, nDCG@k) for the epoch-end hooks.
The reference computes them with torchmetrics (``nrms_module.py:171-187``), which is not
installed here; these are vectorised restatements of the standard definitions and are not
part of the timed hot path."""
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


def ranking_metrics(preds: torch.Tensor, targets: torch.Tensor, sizes: torch.Tensor,
                    top_k_list: List[int]) -> Dict[str, torch.Tensor]:
    """preds/targets: concatenated per-impression scores/labels; sizes: candidates per impression."""
    sizes = sizes.to(preds.device)
    p, t, valid = _dense(preds, targets, sizes)
    order = p.argsort(dim=1, descending=True)
    ts = t.gather(1, order)
    ranks = torch.arange(1, p.shape[1] + 1, device=p.device, dtype=torch.float32)[None, :]
    has_pos = ts.sum(1) > 0
    first = (ts > 0).float().argmax(dim=1).float() + 1
    mrr = torch.where(has_pos, 1.0 / first, torch.zeros_like(first))[has_pos].mean()
    out = {"mrr": mrr}
    disc = 1.0 / torch.log2(ranks + 1)
    ideal = t.sort(dim=1, descending=True).values
    for k in top_k_list:
        dcg = (ts[:, :k] * disc[:, :k]).sum(1)
        idcg = (ideal[:, :k] * disc[:, :k]).sum(1)
        out[f"ndcg@{k}"] = torch.where(idcg > 0, dcg / idcg.clamp_min(1e-12), torch.zeros_like(dcg))[has_pos].mean()
    # global binary AUROC over all candidates (torchmetrics AUROC(task="binary") semantics)
    pf, tf = preds.float(), targets.float()
    o = pf.argsort()
    r = torch.empty_like(pf)
    r[o] = torch.arange(1, pf.numel() + 1, device=pf.device, dtype=torch.float32)
    npos, nneg = tf.sum(), (1 - tf).sum()
    out["auc"] = (r[tf > 0].sum() - npos * (npos + 1) / 2) / (npos * nneg).clamp_min(1)
    return out
