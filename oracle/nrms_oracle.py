"""CPU oracle for the NewsRecLib two-tower hot path (NRMS / NAML).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``newsreclib_b200``) never
routes through this file and fails loudly when its CUDA library is missing.

It is a plain fp32 restatement (explicit matmul / softmax / tanh, no
``nn.MultiheadAttention``, no ``torch_geometric``) of the reference algorithm.
Every function cites the reference file:line it follows (paths are relative to
the reference repo root, andreeaiana/newsreclib @ f29aea8).

Pinning status: the reference ships no golden vectors or known-answer tests for
this path (its only test file does not parse), so the restatement is pinned
against the reference's own ``nn.Module``s imported from ``/root/reference`` in
the build container: ``oracle/make_golden.py`` runs both on the same seeded
inputs, asserts agreement, and writes the fixtures under ``tests/golden/`` that
the ``-m "not gpu"`` suite re-checks the oracle against on every run.
``to_dense_batch`` (torch_geometric 2.3.0, not vendored, not installable here)
is restated from its published algorithm and is "parity unpinned" by the
reference; it is anchored on the reference call sites
(``newsreclib/models/general_rec/nrms_module.py:233,237,277-284``).
The supervised-contrastive loss (``sup_con_loss``) sits on pytorch-metric-learning
2.2.0 (``setup.py:26``; third-party, absent): the reference's own ``losses.py`` and
``model_step`` run unmodified over ``oracle/pml_standins.py`` (a restatement of the
few published pieces they call -- "parity unpinned" for those) and
``oracle/make_module_golden.py`` asserts that ``sup_con_loss`` reproduces them.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# a2  nn.Embedding.from_pretrained(..., freeze=False, padding_idx=0)
# --------------------------------------------------------------------------------------
def embedding_gather(table: Tensor, ids: Tensor) -> Tensor:
    """Row gather, ``text.py:215-217`` (construction) and ``text.py:224`` (forward).

    Row 0 is a real row of the table (``padding_idx`` only zeroes its gradient), so the
    gather is a plain bit-exact ``index_select``.
    """
    return table.index_select(0, ids.reshape(-1)).reshape(*ids.shape, table.shape[1])


# --------------------------------------------------------------------------------------
# a4  nn.MultiheadAttention, math path (SURVEY.md §8 a-notes 1-4)
# --------------------------------------------------------------------------------------
def mha_seq_first(
    x: Tensor, w_in: Tensor, b_in: Tensor, w_out: Tensor, b_out: Tensor, num_heads: int
) -> Tensor:
    """Self-attention over dim 0 of ``x`` ``[L, N, E]`` (``batch_first=False``).

    Follows ``torch.nn.functional.multi_head_attention_forward`` as called by
    ``text.py:229`` and ``encoders/user/nrms.py:34-36``: in-projection with the packed
    ``[3E, E]`` weight (rows 0..E-1 = Q, E..2E-1 = K, 2E..3E-1 = V), heads own contiguous
    feature slices of width ``E / num_heads``, the scale ``1/sqrt(d_h)`` is applied to
    ``q`` before the product, no mask, no attention dropout, then the out-projection.
    """
    L, N, E = x.shape
    d_h = E // num_heads
    qkv = x.reshape(L * N, E) @ w_in.t() + b_in
    q, k, v = qkv[:, :E], qkv[:, E : 2 * E], qkv[:, 2 * E :]
    # [L, N*h, d_h] -> [N*h, L, d_h]; problem index = n*h + head
    q = q.reshape(L, N * num_heads, d_h).transpose(0, 1)
    k = k.reshape(L, N * num_heads, d_h).transpose(0, 1)
    v = v.reshape(L, N * num_heads, d_h).transpose(0, 1)
    q = q * math.sqrt(1.0 / d_h)
    p = torch.softmax(q @ k.transpose(1, 2), dim=-1)
    o = p @ v  # [N*h, L, d_h]
    o = o.transpose(0, 1).reshape(L * N, E)
    y = o @ w_out.t() + b_out
    return y.reshape(L, N, E)


# --------------------------------------------------------------------------------------
# a5  AdditiveAttention
# --------------------------------------------------------------------------------------
def additive_attention(x: Tensor, w: Tensor, b: Tensor, query: Tensor) -> Tensor:
    """``layers/attention.py:34-40``: ``softmax(tanh(xW^T+b)·q, dim=1)``-weighted sum of x."""
    a = torch.tanh(x @ w.t() + b)
    wts = torch.softmax(a @ query, dim=1)
    return (wts.unsqueeze(1) @ x).squeeze(1)


# --------------------------------------------------------------------------------------
# a6  MHSAAddAtt.forward  (title -> news vector)
# --------------------------------------------------------------------------------------
def mhsa_add_att(
    ids: Tensor,
    p: Dict[str, Tensor],
    num_heads: int,
    keep1: Optional[Tensor] = None,
    keep2: Optional[Tensor] = None,
    dropout_p: float = 0.0,
) -> Tensor:
    """``text.py:222-236``.  ``p`` uses the reference ``state_dict`` key names relative to
    the text encoder (``embedding_layer.weight``, ``multihead_attention.in_proj_weight`` …).

    Dropout (``text.py:225,230``) is expressed through explicit keep masks ``[N, L, E]`` so
    a train-mode comparison can feed the very mask the CUDA kernel drew; ``None`` = eval.
    """
    x = embedding_gather(p["embedding_layer.weight"], ids)  # [N, L, E]  text.py:224
    if keep1 is not None:
        x = x * keep1 / (1.0 - dropout_p)  # text.py:225
    x = x.permute(1, 0, 2)  # text.py:228
    y = mha_seq_first(
        x,
        p["multihead_attention.in_proj_weight"],
        p["multihead_attention.in_proj_bias"],
        p["multihead_attention.out_proj.weight"],
        p["multihead_attention.out_proj.bias"],
        num_heads,
    )  # text.py:229
    y = y.permute(1, 0, 2)  # text.py:233
    if keep2 is not None:
        y = y * keep2 / (1.0 - dropout_p)  # text.py:230
    return additive_attention(
        y,
        p["additive_attention.linear.weight"],
        p["additive_attention.linear.bias"],
        p["additive_attention.query"],
    )  # text.py:234


# --------------------------------------------------------------------------------------
# a8  torch_geometric.utils.to_dense_batch (2.3.0)
# --------------------------------------------------------------------------------------
def to_dense_batch(x: Tensor, batch: Tensor) -> Tuple[Tensor, Tensor]:
    """Ragged -> dense.  Restated from torch_geometric 2.3.0 ``utils/to_dense_batch.py``
    (third-party, pinned in the reference's ``setup.py:21`` / ``requirements.txt:28``);
    call sites ``nrms_module.py:233,237,277-284``.  ``batch`` must be sorted."""
    B = int(batch.max()) + 1
    num = torch.zeros(B, dtype=torch.long).scatter_add_(0, batch, torch.ones_like(batch))
    cum = torch.cat([num.new_zeros(1), num.cumsum(0)])
    M = int(num.max())
    idx = torch.arange(batch.numel()) - cum[batch] + batch * M
    out = x.new_zeros((B * M,) + tuple(x.shape[1:]))
    out[idx] = x
    mask = torch.zeros(B * M, dtype=torch.bool)
    mask[idx] = True
    return out.view((B, M) + tuple(x.shape[1:])), mask.view(B, M)


# --------------------------------------------------------------------------------------
# a9  NRMS UserEncoder  (batch-axis attention quirk)
# --------------------------------------------------------------------------------------
def nrms_user_encoder(h: Tensor, p: Dict[str, Tensor], num_heads: int) -> Tensor:
    """``encoders/user/nrms.py:32-41``.  ``h`` is ``[B, Hmax, E]`` and is handed to a
    ``batch_first=False`` attention WITHOUT a permute, so dim 0 (the B impressions) is the
    sequence axis and dim 1 (history position) the batch axis; additive pooling then runs
    over dim 1 (Hmax, zero-padded rows included)."""
    y = mha_seq_first(
        h,
        p["multihead_attention.in_proj_weight"],
        p["multihead_attention.in_proj_bias"],
        p["multihead_attention.out_proj.weight"],
        p["multihead_attention.out_proj.bias"],
        num_heads,
    )
    return additive_attention(
        y,
        p["additive_attention.linear.weight"],
        p["additive_attention.linear.bias"],
        p["additive_attention.query"],
    )


# --------------------------------------------------------------------------------------
# a11 DotProduct, a12 CrossEntropyLoss with soft targets
# --------------------------------------------------------------------------------------
def dot_product(user: Tensor, cand: Tensor) -> Tensor:
    """``layers/click_predictor.py:9-11`` as called at ``nrms_module.py:251-253``:
    ``bmm(user[B,1,E], cand[B,Cmax,E]^T) -> [B, Cmax]``; zero-padded slots score 0.0."""
    return torch.bmm(user.unsqueeze(1), cand.permute(0, 2, 1)).squeeze(1)


def ce_soft(scores: Tensor, y: Tensor) -> Tensor:
    """``nrms_module.py:288`` with ``CrossEntropyLoss()`` (``abstract_recommender.py:115-116``)
    on float (probability) targets: ``mean_b( -sum_c y[b,c] * log_softmax(s[b,:])[c] )``.
    Padded slots (score 0, target 0) take part in the softmax denominator."""
    return -(y * torch.log_softmax(scores, dim=1)).sum(dim=1).mean()


# --------------------------------------------------------------------------------------
# NRMSModule.forward / loss
# --------------------------------------------------------------------------------------
def split_params(params: Dict[str, Tensor]) -> Tuple[Dict[str, Tensor], Dict[str, Tensor]]:
    """Split a reference-named ``NRMSModule.state_dict()`` into the title
    text-encoder dict and the user-encoder dict."""
    tp = "news_encoder.text_encoders.title."
    up = "user_encoder."
    title = {k[len(tp):]: v for k, v in params.items() if k.startswith(tp)}
    user = {k[len(up):]: v for k, v in params.items() if k.startswith(up)}
    return title, user


def nrms_forward(
    batch: Dict, params: Dict[str, Tensor], num_heads: int, late_fusion: bool = False,
    masks: Optional[Dict[str, Tensor]] = None, dropout_p: float = 0.0,
) -> Tensor:
    """``nrms_module.py:230-255``: history and candidate titles through the same news
    encoder, ragged->dense, user encoder (or the late-fusion mean, ``:243-248``), scorer."""
    title, user = split_params(params)
    m = masks or {}
    hist = mhsa_add_att(batch["x_hist"]["title"], title, num_heads,
                        m.get("hist1"), m.get("hist2"), dropout_p)  # :232
    hist_agg, mask_hist = to_dense_batch(hist, batch["batch_hist"])  # :233
    cand = mhsa_add_att(batch["x_cand"]["title"], title, num_heads,
                        m.get("cand1"), m.get("cand2"), dropout_p)  # :236
    cand_agg, _ = to_dense_batch(cand, batch["batch_cand"])  # :237
    if not late_fusion:
        u = nrms_user_encoder(hist_agg, user, num_heads)  # :241
    else:
        hist_size = mask_hist.sum(dim=1)  # :244-247
        u = hist_agg.sum(dim=1) / hist_size.unsqueeze(-1)  # :248
    return dot_product(u, cand_agg)  # :251-253


def sup_con_loss(scores: Tensor, y_true: Tensor, mask_cand: Tensor, temperature: float = 0.1) -> Tensor:
    """Supervised-contrastive loss over the dense score matrix, ``nrms_module.py:290-316`` +
    ``models/components/losses.py:12-40`` on top of pytorch-metric-learning 2.2.0 (``setup.py:26``; third-party, absent
    from /root/reference and from this image: its pieces below are restated from the published source and are "parity
    unpinned" by the reference -- ``oracle/make_module_golden.py`` runs the reference's OWN ``losses.py`` and
    ``model_step`` over ``oracle/pml_standins.py`` and asserts this function reproduces them).

    * index tuples (``nrms_module.py:291-307``): positives of row ``i`` = slots with a non-zero label; negatives = the
      first ``count_i - npos_i`` slots whose label is zero -- real candidates come first in the dense row, so these are
      exactly the real candidates with label 0; padded slots are in neither set;
    * ``GenericPairLoss.compute_loss`` as overridden at ``losses.py:12-17``: zero loss when every index list has at most
      one element; otherwise ``mat`` = the SCORES themselves (no distance), ``pos_mask[a1, p] = 1``,
      ``neg_mask[a2, n] = 1`` (``mat_based_loss``);
    * ``losses.py:19-40``: ``mat / temperature``, minus the detached row maximum (over ALL slots of the row, padded
      zeros included), ``denominator = logsumexp`` over the kept (pos + neg) slots (``lmu.logsumexp``: the others are
      filled with ``finfo.min``), ``-sum(pos * log_prob) / (npos + finfo.tiny)`` per row;
    * reduction: ``SupConLoss.get_default_reducer() = AvgNonZeroReducer`` -- the mean over the rows whose loss is > 0
      (rows without a positive give exactly 0 and are left out), zero when there is none.

    ``abstract_recommender.py:117-120`` builds ``SupConLoss()`` with the DEFAULT temperature 0.1: the module's
    ``temperature`` hyper-parameter never reaches it."""
    pos = (y_true != 0) & mask_cand
    neg = (y_true == 0) & mask_cand
    n_pos, n_neg = int(pos.sum()), int(neg.sum())
    if (n_pos <= 1 and n_neg <= 1) or n_pos == 0 or n_neg == 0:  # losses.py:14-15,20,40 -> zero_losses()
        return (scores * 0).sum()
    mat = scores / temperature
    mat = mat - mat.max(dim=1, keepdim=True)[0].detach()
    keep = pos | neg
    denominator = torch.logsumexp(mat.masked_fill(~keep, torch.finfo(mat.dtype).min), dim=1, keepdim=True)
    denominator = denominator.masked_fill(~keep.any(dim=1, keepdim=True), 0)
    log_prob = mat - denominator
    posf = pos.to(mat.dtype)
    row = -(posf * log_prob).sum(dim=1) / (posf.sum(dim=1) + torch.finfo(mat.dtype).tiny)
    sel = row > 0
    return row[sel].mean() if bool(sel.any()) else (scores * 0).sum()


def nrms_loss(batch: Dict, scores: Tensor, loss: str = "cross_entropy_loss",
              dual_loss_coef: Optional[float] = None) -> Tensor:
    """``nrms_module.py:277,286-328``: soft-target CE, SupCon, or their weighted average (``dual_loss``)."""
    y_true, mask_cand = to_dense_batch(batch["labels"], batch["batch_cand"])
    if loss == "cross_entropy_loss":
        return ce_soft(scores, y_true)
    scl = sup_con_loss(scores, y_true, mask_cand)
    if loss == "sup_con_loss":
        return scl
    assert loss == "dual_loss" and dual_loss_coef is not None
    return (1 - dual_loss_coef) * ce_soft(scores, y_true) + dual_loss_coef * scl


# --------------------------------------------------------------------------------------
# torch.optim.Adam (the reference optimizer, configs/model/nrms.yaml:49-52)
# --------------------------------------------------------------------------------------
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float = 1e-4,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """One in-place ``torch.optim.Adam`` update (no weight decay, no amsgrad), restating
    ``torch/optim/adam.py::_single_tensor_adam`` (third-party, torch 2.x): dense over the
    whole tensor, so rows with zero gradient still see their moments decay."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


# --------------------------------------------------------------------------------------
# a13  NAML pieces
# --------------------------------------------------------------------------------------
def cnn_add_att(ids: Tensor, p: Dict[str, Tensor], window: int, keep1: Optional[Tensor] = None,
                keep2: Optional[Tensor] = None, dropout_p: float = 0.0) -> Tensor:
    """``text.py:163-176``: embedding -> dropout -> Conv2d(1,F,(w,E),pad((w-1)//2,0)) -> ReLU ->
    dropout -> additive pooling over the tokens.  The conv is written as an explicit
    sliding-window contraction; dropout is expressed through keep masks (``keep1`` ``[N, L, E]``
    after the embedding, ``keep2`` ``[N, L, F]`` after the ReLU; ``None`` = eval mode)."""
    x = embedding_gather(p["embedding_layer.weight"], ids)  # [N, L, E]  text.py:165
    if keep1 is not None:
        x = x * keep1 / (1.0 - dropout_p)  # text.py:166
    N, L, E = x.shape
    pad = int((window - 1) / 2)
    w = p["cnn.weight"]  # [F, 1, w, E]
    F_ = w.shape[0]
    xp = torch.zeros(N, L + 2 * pad, E, dtype=x.dtype)
    xp[:, pad : pad + L] = x
    Lout = L + 2 * pad - window + 1
    cols = torch.stack([xp[:, i : i + Lout] for i in range(window)], dim=2)  # [N, Lout, w, E]
    y = cols.reshape(N * Lout, window * E) @ w.reshape(F_, window * E).t() + p["cnn.bias"]  # text.py:169
    y = torch.relu(y).reshape(N, Lout, F_)  # text.py:170
    if keep2 is not None:
        y = y * keep2 / (1.0 - dropout_p)  # text.py:171
    return additive_attention(
        y,
        p["additive_attention.linear.weight"],
        p["additive_attention.linear.bias"],
        p["additive_attention.query"],
    )  # text.py:174


def linear_category_encoder(ids: Tensor, p: Dict[str, Tensor]) -> Tensor:
    """``category.py:73-82`` (eval mode, ``use_dropout``-independent):
    embedding -> Linear -> ReLU."""
    x = embedding_gather(p["embedding_layer.weight"], ids)
    return torch.relu(x @ p["linear.weight"].t() + p["linear.bias"])


def naml_user_encoder(h: Tensor, p: Dict[str, Tensor]) -> Tensor:
    """``encoders/user/naml.py:27-31``: additive pooling over dim 1 only."""
    return additive_attention(
        h,
        p["additive_attention.linear.weight"],
        p["additive_attention.linear.bias"],
        p["additive_attention.query"],
    )


# --------------------------------------------------------------------------------------
# NAMLModule.forward (naml_module.py:261-286) on reference-named parameters
# --------------------------------------------------------------------------------------
def _sub(params: Dict[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    return {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}


def naml_news_encoder(x: Dict[str, Tensor], params: Dict[str, Tensor], window: int,
                      masks: Optional[Dict[str, Tensor]] = None, dropout_p: float = 0.0) -> Tensor:
    """``NewsEncoder.forward`` in the NAML configuration (``news.py:134-163``): the SAME text
    encoder on title and abstract (``news.py:68-77``), the category encoder, then additive
    attention over the stacked views (order-independent).  ``params`` are relative to
    ``news_encoder.``; ``masks`` keys: ``title1/title2/abstract1/abstract2``."""
    tp = _sub(params, "text_encoders.title.")
    cp = _sub(params, "category_encoders.category.")
    lp = _sub(params, "combine_layer.")
    m = masks or {}
    views = [cnn_add_att(x[name], tp, window, m.get(name + "1"), m.get(name + "2"), dropout_p)
             for name in ("title", "abstract") if name in x]
    views.append(linear_category_encoder(x["category"], cp))
    return additive_attention(torch.stack(views, dim=1), lp["linear.weight"], lp["linear.bias"], lp["query"])


def naml_forward(batch: Dict, params: Dict[str, Tensor], window: int, late_fusion: bool = False,
                 masks: Optional[Dict[str, Dict[str, Tensor]]] = None, dropout_p: float = 0.0) -> Tensor:
    """``naml_module.py:261-286``.  ``masks`` = {"hist": {...}, "cand": {...}} keep masks."""
    news = _sub(params, "news_encoder.")
    m = masks or {}
    hist = naml_news_encoder(batch["x_hist"], news, window, m.get("hist"), dropout_p)  # :263
    hist_agg, mask_hist = to_dense_batch(hist, batch["batch_hist"])  # :264
    cand = naml_news_encoder(batch["x_cand"], news, window, m.get("cand"), dropout_p)  # :267
    cand_agg, _ = to_dense_batch(cand, batch["batch_cand"])  # :268
    if not late_fusion:
        u = naml_user_encoder(hist_agg, _sub(params, "user_encoder."))  # :272
    else:
        u = hist_agg.sum(dim=1) / mask_hist.sum(dim=1).unsqueeze(-1)  # :275-279
    return dot_product(u, cand_agg)  # :282-284


# --------------------------------------------------------------------------------------
# a14  PLM head: the part of PLM.forward after the transformer
# --------------------------------------------------------------------------------------
def plm_head(x: Tensor, p: Dict[str, Tensor], num_heads: int, keep1: Optional[Tensor] = None,
             keep2: Optional[Tensor] = None, dropout_p: float = 0.0) -> Tensor:
    """``text.py:93-100``: ``x`` = last hidden states ``[N, T, E]`` -> dropout (``:93``) ->
    ``nn.MultiheadAttention`` called WITHOUT a permute (``:96``; ``batch_first=False``, so dim 0 =
    the N news of the call is the sequence axis and the T tokens are the batch axis) -> dropout
    (``:97``) -> additive pooling over dim 1 = tokens (``:100``, pad tokens included)."""
    if keep1 is not None:
        x = x * keep1 / (1.0 - dropout_p)
    y = mha_seq_first(
        x,
        p["multihead_attention.in_proj_weight"],
        p["multihead_attention.in_proj_bias"],
        p["multihead_attention.out_proj.weight"],
        p["multihead_attention.out_proj.bias"],
        num_heads,
    )
    if keep2 is not None:
        y = y * keep2 / (1.0 - dropout_p)
    return additive_attention(
        y,
        p["additive_attention.linear.weight"],
        p["additive_attention.linear.bias"],
        p["additive_attention.query"],
    )
