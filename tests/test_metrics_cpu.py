"""Epoch-end ranking metrics against independent implementations: scikit-learn for AUROC and
nDCG, a python loop for MRR.  The reference uses torchmetrics (nrms_module.py:182-191), absent from this image;
its conventions (mean over all impressions, an impression without a positive counts 0, mid-rank ties in the
AUROC) are the ones asserted here."""
import numpy as np
import pytest
import torch

from newsreclib_b200.metrics import binary_auroc, ranking_metrics

sk = pytest.importorskip("sklearn.metrics")


def _case(seed, n_imp=40, max_c=30, quantise=False, empty=False):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(2, max_c + 1, n_imp)
    preds, targets = [], []
    for i, c in enumerate(sizes):
        s = rng.normal(size=c).astype(np.float32)
        if quantise:
            s = np.round(s * 2) / 2                      # many tied scores
        y = np.zeros(c, np.float32)
        if not (empty and i % 7 == 0):
            y[rng.choice(c, size=rng.integers(1, max(2, c // 3)), replace=False)] = 1
        preds.append(s); targets.append(y)
    return preds, targets, sizes


@pytest.mark.parametrize("seed,quantise,empty", [(0, False, False), (1, False, True), (2, True, False)])
def test_ranking_metrics_match_independent_implementations(seed, quantise, empty):
    preds, targets, sizes = _case(seed, quantise=quantise, empty=empty)
    got = ranking_metrics(torch.from_numpy(np.concatenate(preds)), torch.from_numpy(np.concatenate(targets)),
                          torch.from_numpy(sizes), [5, 10])
    allp, allt = np.concatenate(preds), np.concatenate(targets)
    assert float(got["auc"]) == pytest.approx(sk.roc_auc_score(allt, allp), abs=1e-6)
    if quantise:
        return  # tied scores: retrieval metrics depend on the tie order (torchmetrics averages, argsort does not)
    mrr, ndcg = [], {5: [], 10: []}
    for s, y in zip(preds, targets):
        if y.sum() == 0:
            mrr.append(0.0)
            for k in ndcg:
                ndcg[k].append(0.0)
            continue
        order = np.argsort(-s, kind="stable")
        mrr.append(1.0 / (1 + int(np.argmax(y[order] > 0))))
        for k in ndcg:
            ndcg[k].append(sk.ndcg_score(y[None, :], s[None, :], k=k))
    assert float(got["mrr"]) == pytest.approx(np.mean(mrr), abs=1e-6)
    for k in ndcg:
        assert float(got[f"ndcg@{k}"]) == pytest.approx(np.mean(ndcg[k]), abs=1e-5)


def test_auroc_edge_cases():
    t = torch.tensor([0.0, 1.0, 0.0, 1.0])
    assert float(binary_auroc(torch.tensor([0.1, 0.9, 0.2, 0.8]), t)) == 1.0
    assert float(binary_auroc(torch.tensor([0.9, 0.1, 0.8, 0.2]), t)) == 0.0
    assert float(binary_auroc(torch.zeros(4), t)) == 0.5            # all tied: chance
    assert float(binary_auroc(torch.tensor([0.3, 0.7]), torch.zeros(2))) == 0.0   # no positive: defined as 0


def test_aspect_metrics_match_reference_functional_golden():
    """tests/golden/metrics_ref.npz = per-impression values of the reference's OWN diversity / personalization functions
    (newsreclib/metrics/functional.py, oracle/make_metrics_golden.py), grouped as metrics/base.py:144-181 groups them."""
    import os
    from newsreclib_b200.metrics import aspect_metrics
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics_ref.npz")))
    t = lambda k: torch.from_numpy(g[k])
    ncat, nsent = [int(x) for x in g["num_classes"]]
    for name, cand, hist, nc in (("categ", "cat", "hcat", ncat), ("sent", "sent", "hsent", nsent)):
        per = aspect_metrics(t("preds"), t("sizes"), t(cand), t(hist), t("hist_sizes"), nc, [5, 10], name, per_impression=True)
        mean = aspect_metrics(t("preds"), t("sizes"), t(cand), t(hist), t("hist_sizes"), nc, [5, 10], name)
        for k in (5, 10):
            for kind in ("div", "pers"):
                key = f"{name}_{kind}@{k}"
                assert np.allclose(per[key].numpy(), g[key], atol=2e-6), key
                assert float(mean[key]) == pytest.approx(float(g[key].mean()), abs=2e-6)
    assert float(g["categ_div@5"][0]) == 0.0 and float(g["categ_pers@5"][0]) == 0.0   # the impression with aspect ids all 0
