"""Per-impression ranking metrics on device (AUC, MRR, nDCG@k) and the aspect-based diversity / personalization of the
test epoch (``newsreclib/metrics/functional.py``) for the epoch-end hooks.

The reference computes them with torchmetrics (``nrms_module.py:182-191``: ``AUROC(task="binary")`` over all
candidates of the epoch, ``RetrievalMRR()`` and ``RetrievalNormalizedDCG(top_k=k)`` with ``indexes`` = the
impression of each candidate), which is not installed here; these are vectorised restatements of those
definitions (third-party, torchmetrics 1.x):

* retrieval metrics are computed per impression and averaged over ALL impressions; an impression without a
  positive counts as 0 (torchmetrics' default ``empty_target_action="neg"``);
* nDCG uses the labels as gains, discount ``1 / log2(rank + 1)``, and the ideal ordering of the same impression;
* AUROC is the global binary AUROC over every candidate, ties in the scores counted half (mid-ranks), which is
  what the trapezoidal ROC integral gives.

On CUDA tensors MRR / nDCG@k come from ``nrl_rank_metrics`` (one CTA per impression, ranks counted in shared memory);
the torch formulation below (one dense ``[impressions, max candidates]`` sort) states the same definitions for host
tensors and is what the CPU tests pin against scikit-learn.  AUROC is global over the epoch: one ``torch.unique``.
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
    if preds.is_cuda and len(top_k_list) <= 4:
        # the sm_100a kernel: one CTA per impression ranks its candidates in shared memory (no dense [B, Cmax] sort)
        from . import ops
        per = ops.rank_metrics(preds, targets, sizes, top_k_list).mean(dim=0)
        out = {"mrr": per[0]}
        out.update({f"ndcg@{k}": per[1 + i] for i, k in enumerate(top_k_list)})
        out["auc"] = binary_auroc(preds, targets)
        return out
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


def aspect_metrics(preds: torch.Tensor, sizes: torch.Tensor, cand_aspects: torch.Tensor, hist_aspects: torch.Tensor,
                   hist_sizes: torch.Tensor, num_classes: int, top_k_list: List[int], name: str,
                   per_impression: bool = False) -> Dict[str, torch.Tensor]:
    """Aspect-based diversity ``D@k`` and personalization ``PS@k`` of the reference's test epoch
    (``metrics/functional.py:8-49`` and ``:52-110`` through ``metrics/diversity.py`` / ``personalization.py``, logged at
    ``nrms_module.py:494-522`` as ``{name}_div@k`` / ``{name}_pers@k`` with name = ``categ`` | ``sent``), for all
    impressions at once:

    * the k best-scored candidates of an impression -> histogram of their aspect ids over ``num_classes`` bins;
    * diversity = entropy of that histogram (normalised to a distribution) / log(num_classes);
    * personalization = generalised Jaccard ``sum(min) / sum(max)`` between that histogram and the histogram of the
      aspects of the impression's clicked history;
    * an impression whose candidate aspect ids sum to 0 counts 0.0 (``empty_target_action="neg"``,
      ``metrics/base.py:160-170``); the result is the mean over all impressions.

    preds / cand_aspects: concatenated per impression (``sizes``); hist_aspects likewise (``hist_sizes``).  One dense
    sort and two scatter-adds; no loop over impressions."""
    dev = preds.device
    sizes, hist_sizes = sizes.to(dev).long(), hist_sizes.to(dev).long()
    B, M = sizes.numel(), int(sizes.max())
    valid = torch.arange(M, device=dev)[None, :] < sizes[:, None]
    p = torch.full((B, M), float("-inf"), device=dev)
    a = torch.zeros(B, M, dtype=torch.long, device=dev)
    p[valid] = preds.float()
    a[valid] = cand_aspects.to(dev).long()
    order = p.argsort(dim=1, descending=True, stable=True)
    a_sorted = a.gather(1, order)
    has_target = a.sum(dim=1) != 0
    hist_rows = torch.repeat_interleave(torch.arange(B, device=dev), hist_sizes)
    hist_count = torch.zeros(B, num_classes, device=dev)
    hist_count.index_put_((hist_rows, hist_aspects.to(dev).long()), torch.ones((), device=dev), accumulate=True)
    log_c = torch.log(torch.tensor(float(num_classes), device=dev))
    zero = torch.zeros(B, device=dev)
    out = {}
    for k in top_k_list:
        kk = min(k, M)
        take = valid[:, :kk].float()                     # ranks beyond the impression's own candidates do not count
        count = torch.zeros(B, num_classes, device=dev).scatter_add_(1, a_sorted[:, :kk], take)
        prob = count / count.sum(dim=1, keepdim=True).clamp_min(1.0)
        ent = -(prob * torch.log(prob.clamp_min(torch.finfo(prob.dtype).tiny))).sum(dim=1)
        div = torch.where(has_target, ent / log_c, zero)
        jac = torch.minimum(count, hist_count).sum(dim=1) / torch.maximum(count, hist_count).sum(dim=1).clamp_min(1.0)
        pers = torch.where(has_target, jac, zero)
        out[f"{name}_div@{k}"] = div if per_impression else div.mean()
        out[f"{name}_pers@{k}"] = pers if per_impression else pers.mean()
    return out
