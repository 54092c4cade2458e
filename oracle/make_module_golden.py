"""Pin the oracle's glue against the reference's OWN ``NRMSModule`` (``forward`` + ``model_step``) and mint
``tests/golden/nrms_module_ref_*.npz``.

Runs ONLY in the build container (imports the unmodified reference from ``/root/reference``).
``newsreclib/models/general_rec/nrms_module.py`` needs ``lightning``, ``torchmetrics``, ``torch_geometric`` and (through
``newsreclib.models.components.losses``) ``pytorch_metric_learning``; none is installed and none takes part in the
arithmetic of ``forward`` / ``model_step`` except ``torch_geometric.utils.to_dense_batch``.  The imports are satisfied
with stand-ins (``oracle/ref_standins.py``):

* ``lightning.LightningModule`` -> ``torch.nn.Module`` + ``save_hyperparameters`` (constructor arguments -> ``self.hparams``)
  and a ``device`` property;
* ``torchmetrics`` metric classes, ``newsreclib.metrics.*`` and ``SupConLoss`` -> inert objects (built in ``__init__``,
  never called by ``forward`` / ``model_step`` with the cross-entropy loss);
* ``torch_geometric.utils.to_dense_batch`` -> ``oracle.nrms_oracle.to_dense_batch``, the restatement of the published
  PyG 2.3.0 algorithm -- third-party, absent, "parity unpinned" for that one function.

``nrms_module.py`` and ``abstract_recommender.py`` themselves are loaded UNMODIFIED, so the scores, the loss and the
11-tuple below come out of the reference's own ``NRMSModule.forward`` (``:230-255``) and ``model_step`` (``:260-362``).
The script asserts that ``oracle/nrms_oracle.py`` reproduces them and stores batch, reference outputs and gradients.

Usage:  python oracle/make_module_golden.py [--check]
"""
from __future__ import annotations

import functools
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import nrms_oracle as O  # noqa: E402
from oracle._golden_io import save as golden_save  # noqa: E402


# stand-ins for the absent packages (shared with bench.py's reference arm): oracle/ref_standins.py
from oracle import ref_standins  # noqa: E402

ref_standins.install("/root/reference")

from newsreclib.models.general_rec.nrms_module import NRMSModule  # noqa: E402  (the reference's own file)

from newsreclib_b200.synthetic import make_batch, make_nrms_params  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
OUTPUTS = {"train": ["preds", "targets", "cand_news_size"], "val": ["preds", "targets", "cand_news_size"],
           "test": ["preds", "targets", "cand_news_size", "hist_news_size", "user_ids", "cand_news_ids"]}
TITLE = "news_encoder.text_encoders.title."
GRAD_STRIDE = 37


def build_reference(params, late_fusion, tmpdir, loss="cross_entropy_loss", dual_loss_coef=None):
    emb = os.path.join(tmpdir, "emb.npy")
    np.save(emb, params[TITLE + "embedding_layer.weight"].numpy())
    m = NRMSModule(
        dataset_attributes=["title", "category"], attributes2encode=["title"], outputs=OUTPUTS,
        dual_loss_training=loss == "dual_loss", dual_loss_coef=dual_loss_coef, loss=loss, late_fusion=late_fusion,
        temperature=0.36 if loss != "cross_entropy_loss" else None,  # never reaches SupConLoss() (abstract_recommender.py:117-120)
        use_plm=False, pretrained_embeddings_path=emb, plm_model=None, frozen_layers=None,
        embed_dim=300, num_heads=15, query_dim=200, dropout_probability=0.2, top_k_list=[5, 10],
        num_categ_classes=18, num_sent_classes=3, save_recs=False, recs_fpath=None,
        optimizer=functools.partial(torch.optim.Adam, lr=1e-4), scheduler=None)
    own = {k: v for k, v in params.items() if k in m.state_dict()}
    missing = m.load_state_dict(own, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.eval()


def rel(a, b):
    a, b = a.detach(), b.detach()
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def mint(name, V, B, max_hist, seed, late_fusion, criterion="cross_entropy_loss", dual_loss_coef=None, cand="train"):
    params = make_nrms_params(V, seed=seed)
    batch = make_batch(B, V, hist="ragged", cand=cand, seed=seed, max_hist=max_hist)
    if criterion != "cross_entropy_loss":
        # the SupCon index tuples (nrms_module.py:290-307) on rows the plain generator does not make: a row with
        # several positives, a row with none (its loss is exactly 0 and AvgNonZeroReducer leaves it out), ragged
        # candidate counts (padded slots are in neither the positive nor the negative set)
        off = np.concatenate([[0], np.cumsum(np.bincount(batch["batch_cand"].numpy(), minlength=B))])
        batch["labels"][off[1]:off[1] + 2] = 1.0
        batch["labels"][off[1] + 2 + (off[2] - off[1] - 2) // 2] = 1.0
        batch["labels"][off[2]:off[3]] = 0.0
    with tempfile.TemporaryDirectory() as tmp:
        m = build_reference(params, late_fusion, tmp, criterion, dual_loss_coef)
    scores = m(batch)                                   # NRMSModule.forward, nrms_module.py:230-255
    out = m.model_step(batch)                           # nrms_module.py:260-362
    assert len(out) == 11
    loss = out[0]
    loss.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in m.named_parameters()}
    # the oracle's restatement of the same glue
    ps = {k: v.clone().requires_grad_(True) for k, v in params.items() if k in grads}
    o_scores = O.nrms_forward(batch, ps, 15, late_fusion=late_fusion)
    o_loss = O.nrms_loss(batch, o_scores, criterion, dual_loss_coef)
    o_loss.backward()
    assert o_scores.shape == scores.shape and rel(o_scores, scores) < 2e-6, rel(o_scores, scores)
    assert rel(o_loss, loss) < 2e-6
    # arbiter for the gradients: the same oracle in float64.  A few gradients (the additive-attention bias: sum_t ds_t = 0
    # per softmax group) cancel to ~1e-5 of their terms, so two fp32 evaluations agree only to their own noise there.
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in params.items() if k in grads}
    b64 = dict(batch)
    b64["labels"] = batch["labels"].double()
    O.nrms_loss(b64, O.nrms_forward(b64, p64, 15, late_fusion=late_fusion), criterion, dual_loss_coef).backward()
    for k, g in grads.items():
        og = ps[k].grad.clone() if ps[k].grad is not None else torch.zeros_like(ps[k])
        g64 = p64[k].grad.clone() if p64[k].grad is not None else torch.zeros_like(p64[k])
        if k == TITLE + "embedding_layer.weight":
            og[0] = 0  # nn.Embedding(padding_idx=0) never accumulates into row 0 (text.py:215-217), as helpers.oracle_run does
            g64[0] = 0
            assert float(g[0].abs().max()) == 0.0
        if float(g64.abs().max()) < 1e-9:
            continue  # mathematically zero (the key bias of in_proj_bias): both fp32 runs hold rounding noise
        ref_noise = rel(g, g64)
        assert ref_noise <= 2e-3, (k, ref_noise)                           # the reference agrees with the fp64 oracle
        assert rel(og, g64) <= max(2e-5, 4.0 * ref_noise) + 1e-12, (k, rel(og, g64), ref_noise)
    names = ["loss", "preds", "targets", "cand_news_size", "hist_news_size", "target_categories", "target_sentiments",
             "hist_categories", "hist_sentiments", "user_ids", "cand_news_ids"]
    rec = {"meta": np.array([300, 15, 200, V, B, max_hist, seed, 30, int(late_fusion)]), "scores": scores.detach().numpy()}
    if criterion != "cross_entropy_loss":
        rec["loss_name"] = np.array(criterion)
        rec["dual_loss_coef"] = np.array(0.0 if dual_loss_coef is None else dual_loss_coef)
        s_leaf = scores.detach().clone().requires_grad_(True)   # d loss / d scores of the reference's own criterion
        O_y, O_mask = O.to_dense_batch(batch["labels"], batch["batch_cand"])
        crit = (lambda s: m.criterion(embeddings=s, labels=None, indices_tuple=_indices(O_y, O_mask), ref_emb=None,
                                      ref_labels=None)) if criterion == "sup_con_loss" else (
            lambda s: (1 - dual_loss_coef) * m.ce_criterion(s, O_y) + dual_loss_coef * m.scl_criterion(
                embeddings=s, labels=None, indices_tuple=_indices(O_y, O_mask), ref_emb=None, ref_labels=None))
        l2 = crit(s_leaf)
        assert rel(l2, out[0]) < 1e-6, (float(l2), float(out[0]))
        l2.backward()
        rec["d_scores"] = s_leaf.grad.numpy()
    for n, t in zip(names, out):
        rec["out/" + n] = t.detach().numpy()
    for k, g in grads.items():  # small tensors whole, big ones as every GRAD_STRIDE-th element (fixtures stay small)
        rec["grad/" + k] = (g if g.numel() <= 5000 else g.reshape(-1)[::GRAD_STRIDE]).numpy()
    for side in ("x_hist", "x_cand"):
        for k, v in batch[side].items():
            rec[f"batch/{side}/{k}"] = v.numpy()
    for k in ("batch_hist", "batch_cand", "labels", "user_ids", "user_idx"):
        rec["batch/" + k] = batch[k].numpy()
    rec["param_checksum"] = np.array([float(v.double().sum()) for v in params.values()])
    path = os.path.join(GOLD, name + ".npz")
    golden_save(path, **rec)
    print(f"[{name}] reference NRMSModule.forward/model_step == oracle: scores rel {rel(o_scores, scores):.1e}, "
          f"loss rel {rel(o_loss, loss):.1e}; {os.path.getsize(path)} bytes")


def _indices(y_true, mask_cand):
    """The index tuples exactly as ``nrms_module.py:290-307`` builds them (used only to take d loss / d scores of the
    reference's own criterion objects on the reference's own scores)."""
    pos_idx = [torch.where(y_true[i])[0] for i in range(mask_cand.shape[0])]
    pos_repeats = torch.tensor([len(pos_idx[i]) for i in range(len(pos_idx))])
    q_p = torch.repeat_interleave(torch.arange(mask_cand.shape[0]), pos_repeats)
    p = torch.cat(pos_idx)
    neg_idx = [torch.where(~y_true[i].bool())[0][: len(torch.where(mask_cand[i])[0]) - pos_repeats[i]]
               for i in range(mask_cand.shape[0])]
    neg_repeats = torch.tensor([len(t) for t in neg_idx])
    q_n = torch.repeat_interleave(torch.arange(mask_cand.shape[0]), neg_repeats)
    return q_p, p, q_n, torch.cat(neg_idx)


def naml_check(name="naml_mind"):
    """The NAML fixtures (``oracle/make_golden.py::naml_step_case``) were minted from the reference's component modules
    glued by a restatement of ``NAMLModule.forward``.  Here the reference's OWN ``NAMLModule`` (``naml_module.py``,
    loaded unmodified under the stand-ins above) runs ``forward`` / ``model_step`` on the very inputs of such a fixture and
    must reproduce its scores and loss; the module-level outputs go to ``tests/golden/naml_module_ref.npz``."""
    from newsreclib.models.general_rec.naml_module import NAMLModule
    from newsreclib_b200.synthetic import make_naml_params
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    V, E, F_, W, Q, CE, C, B, max_hist, seed, L, LA = [int(x) for x in g["meta"]]
    params = make_naml_params(V, E, F_, W, Q, CE, C, seed=seed)
    assert np.allclose(np.array([float(v.double().sum()) for v in params.values()]), g["param_checksum"], rtol=1e-9)
    batch = {"batch_hist": torch.from_numpy(g["batch_hist"]), "batch_cand": torch.from_numpy(g["batch_cand"]),
             "labels": torch.from_numpy(g["labels"]), "x_hist": {}, "x_cand": {},
             "user_ids": torch.arange(B) + 7, "user_idx": torch.arange(B)}
    for side in ("hist", "cand"):
        n = g[f"{side}_title"].shape[0]
        for attr in ("title", "abstract", "category"):
            batch["x_" + side][attr] = torch.from_numpy(g[f"{side}_{attr}"])
        batch["x_" + side]["sentiment"] = torch.ones(n, dtype=torch.long)
        batch["x_" + side]["news_ids"] = torch.arange(n)
    with tempfile.TemporaryDirectory() as tmp:
        emb = os.path.join(tmp, "emb.npy")
        np.save(emb, params["news_encoder.text_encoders.title.embedding_layer.weight"].numpy())
        m = NAMLModule(
            dataset_attributes=["title", "abstract", "category"], attributes2encode=["title", "abstract", "category"],
            outputs=OUTPUTS, dual_loss_training=False, dual_loss_coef=None, loss="cross_entropy_loss", late_fusion=False,
            temperature=None, use_plm=False, pretrained_embeddings_path=emb, plm_model=None, frozen_layers=None,
            text_embed_dim=E, num_heads=15, num_filters=F_, window_size=W, query_dim=Q, categ_embed_dim=CE,
            dropout_probability=0.2, top_k_list=[5, 10], num_categ_classes=C - 1, num_sent_classes=3, save_recs=False,
            recs_fpath=None, optimizer=functools.partial(torch.optim.Adam, lr=1e-4), scheduler=None)
    full = dict(params)
    for k in list(params):
        if ".text_encoders.title." in k:
            full[k.replace(".title.", ".abstract.")] = params[k]
    res = m.load_state_dict(full, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    m.eval()
    scores = m(batch)
    out = m.model_step(batch)
    assert rel(scores, torch.from_numpy(g["scores"])) < 1e-6 and rel(out[0], torch.from_numpy(g["loss"])) < 1e-6
    golden_save(os.path.join(GOLD, "naml_module_ref.npz"), source=np.array(name), scores=scores.detach().numpy(),
                        loss=out[0].detach().numpy(), preds=out[1].detach().numpy(), targets=out[2].numpy(),
                        cand_news_size=out[3].numpy(), hist_news_size=out[4].numpy())
    print(f"[naml_module_ref] reference NAMLModule.forward/model_step reproduces {name}.npz: scores rel "
          f"{rel(scores, torch.from_numpy(g['scores'])):.1e}, loss rel {rel(out[0], torch.from_numpy(g['loss'])):.1e}")


if __name__ == "__main__":
    torch.manual_seed(0)
    mint("nrms_module_ref", V=500, B=6, max_hist=9, seed=31, late_fusion=False)
    mint("nrms_module_ref_late_fusion", V=400, B=5, max_hist=7, seed=32, late_fusion=True)
    # nrms_module.py:289-328 with the reference's own losses.py (over oracle/pml_standins.py)
    mint("nrms_module_ref_supcon", V=400, B=6, max_hist=6, seed=33, late_fusion=False, criterion="sup_con_loss", cand="eval")
    mint("nrms_module_ref_dual", V=400, B=6, max_hist=6, seed=34, late_fusion=False, criterion="dual_loss",
         dual_loss_coef=0.3, cand="eval")
    naml_check()
