"""Shared helpers for the parity tests (test infrastructure; may import the oracle)."""
import os

import numpy as np
import torch

from newsreclib_b200.synthetic import make_batch, make_nrms_params

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TITLE = "news_encoder.text_encoders.title."
USER = "user_encoder."


def load_golden(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    E, H, Q, V, B, max_hist, seed, L = [int(x) for x in g["meta"]]
    if any(k.startswith("param/") for k in g):
        params = {k[len("param/"):]: torch.from_numpy(g[k]) for k in g if k.startswith("param/")}
    else:
        params = make_nrms_params(V, E, H, Q, seed=seed)
        chk = np.array([float(v.double().sum()) for v in params.values()])
        assert np.allclose(chk, g["param_checksum"], rtol=1e-9), "seeded parameter generator drifted"
    batch = {
        "x_hist": {"title": torch.from_numpy(g["hist_title"])},
        "x_cand": {"title": torch.from_numpy(g["cand_title"])},
        "batch_hist": torch.from_numpy(g["batch_hist"]),
        "batch_cand": torch.from_numpy(g["batch_cand"]),
        "labels": torch.from_numpy(g["labels"]),
    }
    return g, params, batch, dict(E=E, H=H, Q=Q, V=V, B=B, L=L)


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def batch_sizes(batch):
    B = int(batch["batch_hist"].max()) + 1
    Hmax = int(torch.bincount(batch["batch_hist"], minlength=B).max())
    Cmax = int(torch.bincount(batch["batch_cand"], minlength=B).max())
    return B, Hmax, Cmax


def to_dev(batch, dev="cuda"):
    return {
        "x_hist": {"title": batch["x_hist"]["title"].to(dev)},
        "x_cand": {"title": batch["x_cand"]["title"].to(dev)},
        "batch_hist": batch["batch_hist"].to(dev),
        "batch_cand": batch["batch_cand"].to(dev),
        "labels": batch["labels"].to(dev),
    }


def oracle_run(params, batch, H, masks=None, dropout_p=0.0, late_fusion=False, grad=True, loss_name="cross_entropy_loss",
               dual_loss_coef=None):
    """Oracle forward (+ autograd backward) on CPU; returns scores, loss, grads."""
    from oracle import nrms_oracle as O

    ps = {k: v.clone().requires_grad_(grad) for k, v in params.items()}
    scores = O.nrms_forward(batch, ps, H, late_fusion=late_fusion, masks=masks, dropout_p=dropout_p)
    loss = O.nrms_loss(batch, scores, loss_name, dual_loss_coef)
    grads = None
    if grad:
        loss.backward()
        grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v)) for k, v in ps.items()}
        grads[TITLE + "embedding_layer.weight"][0] = 0  # padding_idx=0 (text.py:215-217)
    return scores.detach(), loss.detach(), grads


def gpu_run(params, batch, H, precision=0, do_backward=True, dropout_p=0.0, training=False, seed=0,
            late_fusion=False):
    """The CUDA path through the C ABI (nrl_nrms_step)."""
    from newsreclib_b200 import ops

    dev = "cuda"
    P = {k: v.to(dev).contiguous() for k, v in params.items()}
    b = to_dev(batch, dev)
    B, Hmax, Cmax = batch_sizes(batch)
    table = P[TITLE + "embedding_layer.weight"]
    E = table.shape[1]
    Q = P[TITLE + "additive_attention.query"].numel()
    nb = ops.block_from_dict(P, TITLE)
    ub = ops.block_from_dict(P, USER)
    grads = None
    if do_backward:
        grads = ([torch.zeros_like(t) for t in nb], [torch.zeros_like(t) for t in ub], torch.zeros_like(table))
    scores, loss, _ = ops.nrms_step(b, table, nb, None if late_fusion else ub, ops.dims_of(E, H, Q), B=B,
                                    Hmax=Hmax, Cmax=Cmax, late_fusion=late_fusion, dropout_p=dropout_p,
                                    training=training, seed=seed, grads=grads, precision=precision)
    torch.cuda.synchronize()
    out_g = None
    if do_backward:
        out_g = {TITLE + "embedding_layer.weight": grads[2].cpu()}
        for k, t in zip(ops.BLOCK_KEYS, grads[0]):
            out_g[TITLE + k] = t.cpu()
        for k, t in zip(ops.BLOCK_KEYS, grads[1]):
            out_g[USER + k] = t.cpu()
    return scores.cpu(), loss.cpu().reshape(()), out_g


def grad_tolerances(params, batch, H, base_tol, ref_grads=None, **kw):
    """Per-tensor gradient tolerance = max(base_tol, 4 x the fp32 oracle's own error against an
    fp64 run of the same oracle).  A few gradients (the additive-attention bias: sum_t ds_t = 0
    per softmax group) are sums that cancel to ~1e-5 of their terms, so the reference's fp32
    result is itself only good to a few 1e-4 there; the bar cannot be tighter than its noise."""
    p64 = {k: v.double() for k, v in params.items()}
    b64 = dict(batch)
    b64["labels"] = batch["labels"].double()
    if kw.get("masks"):
        kw = dict(kw)
        kw["masks"] = {k: v.double() for k, v in kw["masks"].items()}
    _, _, g64 = oracle_run(p64, b64, H, **kw)
    if ref_grads is None:
        _, _, ref_grads = oracle_run(params, batch, H, **kw)
    return {k: max(base_tol, 4.0 * rel_err(ref_grads[k], g64[k])) for k in ref_grads}


def report_step(tag, scores, loss, grads, ref_scores, ref_loss, ref_grads, tols, logit_tol, loss_tol):
    """Compare one step with the oracle, PRINT every tensor's error next to the tolerance applied, then assert."""
    es, el = rel_err(scores, ref_scores), rel_err(loss, ref_loss)
    print(f"{tag}: logits rel {es:.2e} (tol {logit_tol:.0e})  loss rel {el:.2e} (tol {loss_tol:.0e})")
    rows, worst = [], 0.0
    for k, g in ref_grads.items():
        e, t = rel_err(grads[k], g), tols[k]
        rows.append((k, e, t))
        worst = max(worst, e / t)
        print(f"  grad {k:<66s} err {e:.2e}  tol {t:.2e}")
    print(f"{tag}: worst gradient error / tolerance {worst:.2f}")
    assert es <= logit_tol and el <= loss_tol, (tag, es, el)
    for k, e, t in rows:
        assert e <= t, (tag, k, e, t)


def load_collate_golden():
    """``tests/golden/collate_ref.npz`` (minted by ``oracle/make_collate_golden.py`` from the reference's own
    ``DatasetCollate``): returns (news columns as lists, {split: (samples, reference batch dict)}, (L_title, L_abs))."""
    g = dict(np.load(os.path.join(GOLD, "collate_ref.npz")))
    L_title, L_abs = int(g["meta"][0]), int(g["meta"][1])
    news = {}
    for k in ("nid", "category_class", "subcategory_class", "sentiment_class", "sentiment_score"):
        news[k] = g[f"news.{k}"].tolist()
    for k in ("tokenized_title", "tokenized_abstract"):
        flat, lens = g[f"news.{k}.flat"], g[f"news.{k}.len"]
        cuts = np.concatenate([[0], np.cumsum(lens)])
        news[k] = [flat[a:b].tolist() for a, b in zip(cuts[:-1], cuts[1:])]
    splits = {}
    for split in ("test", "train"):
        def ragged(name, lens):
            cuts = np.concatenate([[0], np.cumsum(g[f"{split}.{lens}"])])
            return [g[f"{split}.{name}"][a:b] for a, b in zip(cuts[:-1], cuts[1:])]
        hist, cand = ragged("hist_rows", "hist_len"), ragged("cand_rows", "cand_len")
        labels = ragged("labels", "cand_len")
        samples = [(g[f"{split}.user_ids"][i:i + 1], g[f"{split}.user_idx"][i:i + 1], hist[i], cand[i], labels[i])
                   for i in range(len(hist))]
        ref = {}
        for k, v in g.items():
            if k.startswith(f"{split}.batch."):
                name = k[len(f"{split}.batch."):]
                if "." in name:
                    side, col = name.split(".", 1)
                    ref.setdefault(side, {})[col] = torch.from_numpy(v)
                else:
                    ref[name] = torch.from_numpy(v)
        splits[split] = (samples, ref)
    return news, splits, (L_title, L_abs)


MODULE_OUT = ["loss", "preds", "targets", "cand_news_size", "hist_news_size", "target_categories", "target_sentiments",
              "hist_categories", "hist_sentiments", "user_ids", "cand_news_ids"]
GRAD_STRIDE = 37  # oracle/make_module_golden.py stores big gradients as every 37th element


def load_module_golden(name):
    """``tests/golden/nrms_module_ref*.npz``: batch + outputs of the reference's OWN ``NRMSModule.forward`` /
    ``model_step`` (minted by ``oracle/make_module_golden.py``).  Returns (params, batch, reference dict, meta)."""
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    E, H, Q, V, B, max_hist, seed, L, late = [int(x) for x in g["meta"]]
    params = make_nrms_params(V, E, H, Q, seed=seed)
    chk = np.array([float(v.double().sum()) for v in params.values()])
    assert np.allclose(chk, g["param_checksum"], rtol=1e-9), "seeded parameter generator drifted"
    batch = {"x_hist": {}, "x_cand": {}}
    for k, v in g.items():
        if k.startswith("batch/"):
            parts = k.split("/")
            if len(parts) == 3:
                batch[parts[1]][parts[2]] = torch.from_numpy(v)
            else:
                batch[parts[1]] = torch.from_numpy(v)
    ref = {"scores": torch.from_numpy(g["scores"]),
           "out": {n: torch.from_numpy(np.asarray(g["out/" + n])) for n in MODULE_OUT},
           "grad": {k[len("grad/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("grad/")}}
    if "d_scores" in g:  # SupCon / dual-loss fixtures: the criterion's name, its coefficient, d loss / d scores
        ref["loss_name"], ref["dual_loss_coef"] = str(g["loss_name"]), float(g["dual_loss_coef"])
        ref["d_scores"] = torch.from_numpy(g["d_scores"])
    return params, batch, ref, dict(E=E, H=H, Q=Q, V=V, B=B, L=L, late_fusion=bool(late))


def grad_sample(g):
    """The part of a gradient tensor the module fixtures store."""
    g = g.detach().cpu()
    return g if g.numel() <= 5000 else g.reshape(-1)[::GRAD_STRIDE]
