"""GPU parity tests of the NAML rows and the PLM head (a14): the CUDA path through
the C ABI / drop-in modules against the CPU oracle and the golden fixtures minted from the reference's
own modules (oracle/make_golden.py).  Tolerances: logits / vectors 1e-4 relative to max|ref|, loss
1e-4, gradients 1e-3 (or 4x the fp32 oracle's own noise against fp64, whichever is larger)."""
import functools
import os

import numpy as np
import pytest
import torch

from helpers import GOLD, rel_err
from newsreclib_b200.synthetic import make_batch, make_naml_params
from oracle import nrms_oracle as O

pytestmark = pytest.mark.gpu

T_ = "news_encoder.text_encoders.title."
C_ = "news_encoder.category_encoders.category."
OUTPUTS = {"train": ["preds", "targets", "cand_news_size"], "val": ["preds", "targets", "cand_news_size"],
           "test": ["preds", "targets", "cand_news_size", "hist_news_size", "user_ids", "cand_news_ids"]}


def sub(params, prefix):
    return {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}


def oracle_grads(fn, tensors):
    """fn(dict of leaf tensors) -> scalar; returns grads dict (fp32 and fp64 noise estimate)."""
    out = {}
    for dt in (torch.float32, torch.float64):
        leaves = {k: v.detach().to(dt).requires_grad_(True) for k, v in tensors.items()}
        fn(leaves, dt).backward()
        out[dt] = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    tol = {k: max(1e-3, 4.0 * rel_err(out[torch.float32][k], out[torch.float64][k])) for k in tensors}
    return out[torch.float32], tol


# ------------------------------------------------------------------------------------------------
# CNNAddAtt
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dims", [(60, 32, 48, 3, 24, 12, 9), (400, 300, 400, 3, 200, 30, 21), (90, 64, 80, 5, 40, 50, 6)])
def test_cnn_encoder_forward_backward(dims):
    from newsreclib_b200 import ops
    V, E, F_, W, Q, L, n = dims
    p = sub(make_naml_params(V, E, F_, W, Q, 20, 19, seed=V), T_)
    rng = np.random.default_rng(V)
    from newsreclib_b200.synthetic import make_titles
    ids = torch.from_numpy(make_titles(rng, n, V, L, mean_len=L * 0.6, min_len=0))
    wgt = torch.randn(n, F_, generator=torch.Generator().manual_seed(1))
    keys = ["embedding_layer.weight", "cnn.weight", "cnn.bias", "additive_attention.linear.weight",
            "additive_attention.linear.bias", "additive_attention.query"]
    dev = {k: p[k].cuda().requires_grad_(True) for k in keys}
    out = ops.CnnEncoderFn.apply(ids.cuda(), *[dev[k] for k in keys], W, 0.0, False, 0, ops.PREC_BF16X3)
    ref = O.cnn_add_att(ids, p, W)
    assert rel_err(out, ref) <= 1e-4
    (out * wgt.cuda()).sum().backward()

    def f(leaves, dt):
        return (O.cnn_add_att(ids, leaves, W) * wgt.to(dt)).sum()
    rg, tol = oracle_grads(f, {k: p[k] for k in keys})
    rg["embedding_layer.weight"][0] = 0  # padding_idx (text.py:151-153)
    for k in keys:
        assert rel_err(dev[k].grad, rg[k]) <= tol[k], k
    assert float(dev["embedding_layer.weight"].grad[0].abs().max()) == 0.0


def conv_preactivations(ids, p, W, k1, pd):
    """fp64 pre-activations of the conv (oracle math, text.py:169) for the ReLU-kink check."""
    x = O.embedding_gather(p["embedding_layer.weight"].double(), ids) * k1.double() / (1 - pd)
    n, L, E = x.shape
    pad = (W - 1) // 2
    xp = torch.zeros(n, L + 2 * pad, E, dtype=torch.float64)
    xp[:, pad:pad + L] = x
    cols = torch.stack([xp[:, i:i + L] for i in range(W)], dim=2).reshape(n * L, W * E)
    return cols @ p["cnn.weight"].double().reshape(-1, W * E).t() + p["cnn.bias"].double()


def test_cnn_encoder_train_mode_matches_oracle_with_same_mask():
    from newsreclib_b200 import ops
    V, E, F_, W, Q, L, n, pd = 200, 64, 96, 3, 40, 20, 15, 0.2
    p = sub(make_naml_params(V, E, F_, W, Q, 20, 19, seed=3), T_)
    from newsreclib_b200.synthetic import make_titles
    ids = torch.from_numpy(make_titles(np.random.default_rng(1), n, V, L))
    keys = ["embedding_layer.weight", "cnn.weight", "cnn.bias", "additive_attention.linear.weight",
            "additive_attention.linear.bias", "additive_attention.query"]
    # The gradient of a ReLU is discontinuous at 0: a pre-activation within the fp32 noise of 0 makes
    # the comparison ill-posed (the two sides may legitimately pick different branches).  Draw the
    # dropout seed until every kept pre-activation is at least 1e-5 away from the kink.
    for seed in range(4242, 4262):
        k1 = ops.dropout_mask(n * L * E, seed, 0, pd, "cuda").cpu().float().reshape(n, L, E)
        k2 = ops.dropout_mask(n * L * F_, seed, 1, pd, "cuda").cpu().float().reshape(n, L, F_)
        pre = conv_preactivations(ids, p, W, k1, pd)
        if float(pre.abs()[k2.reshape(-1, F_) > 0].min()) > 1e-5:
            break
    else:
        pytest.fail("no well-conditioned dropout seed found")
    assert 0.75 < float(k1.mean()) < 0.85 and 0.75 < float(k2.mean()) < 0.85
    dev = {k: p[k].cuda().requires_grad_(True) for k in keys}
    out = ops.CnnEncoderFn.apply(ids.cuda(), *[dev[k] for k in keys], W, pd, True, seed, ops.PREC_BF16X3)
    ref = O.cnn_add_att(ids, p, W, k1, k2, pd)
    assert rel_err(out, ref) <= 1e-4
    wgt = torch.randn(n, F_, generator=torch.Generator().manual_seed(2))
    (out * wgt.cuda()).sum().backward()

    def f(leaves, dt):
        return (O.cnn_add_att(ids, leaves, W, k1.to(dt), k2.to(dt), pd) * wgt.to(dt)).sum()
    rg, tol = oracle_grads(f, {k: p[k] for k in keys})
    rg["embedding_layer.weight"][0] = 0
    for k in keys:
        assert rel_err(dev[k].grad, rg[k]) <= tol[k], k


# ------------------------------------------------------------------------------------------------
# LinearEncoder (category), AdditiveAttention fwd + bwd
# ------------------------------------------------------------------------------------------------
def test_linear_encoder_forward_backward():
    from newsreclib_b200 import ops
    p = sub(make_naml_params(50, 32, 400, 3, 24, 100, 19, seed=9), C_)
    ids = torch.from_numpy(np.random.default_rng(2).integers(0, 19, 37).astype(np.int64))  # includes id 0
    keys = ["embedding_layer.weight", "linear.weight", "linear.bias"]
    dev = {k: p[k].cuda().requires_grad_(True) for k in keys}
    out = ops.LinearEncoderFn.apply(ids.cuda(), *[dev[k] for k in keys], 0.0, False, 0, ops.PREC_BF16X3)
    ref = O.linear_category_encoder(ids, p)
    assert rel_err(out, ref) <= 1e-4
    assert torch.all(out >= 0)
    wgt = torch.randn(37, 400, generator=torch.Generator().manual_seed(1))
    (out * wgt.cuda()).sum().backward()

    def f(leaves, dt):
        return (O.linear_category_encoder(ids, leaves) * wgt.to(dt)).sum()
    rg, tol = oracle_grads(f, {k: p[k] for k in keys})
    rg["embedding_layer.weight"][0] = 0
    for k in keys:
        assert rel_err(dev[k].grad, rg[k]) <= tol[k], k
    assert float(dev["embedding_layer.weight"].grad[0].abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(7, 3, 400, 200), (5, 50, 400, 200), (64, 11, 48, 24)])
def test_additive_attention_forward_backward(shape):
    from newsreclib_b200.models.components.layers.attention import AdditiveAttention
    G, L, D, Q = shape
    torch.manual_seed(G)
    m = AdditiveAttention(D, Q)
    x = torch.randn(G, L, D)
    x[0, L - 1] = 0  # a zero-padded row takes part in the softmax (no mask in the reference)
    wgt = torch.randn(G, D)
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    mc = m.cuda()
    xc = x.cuda().requires_grad_(True)
    out = mc(xc)
    ref = O.additive_attention(x, p["linear.weight"], p["linear.bias"], p["query"])
    assert rel_err(out, ref) <= 1e-4
    (out * wgt.cuda()).sum().backward()

    def f(leaves, dt):
        return (O.additive_attention(leaves["x"], leaves["linear.weight"], leaves["linear.bias"], leaves["query"])
                * wgt.to(dt)).sum()
    rg, tol = oracle_grads(f, {"x": x, **p})
    assert rel_err(xc.grad, rg["x"]) <= tol["x"]
    for k, v in mc.named_parameters():
        assert rel_err(v.grad, rg[k]) <= tol[k], k


# ------------------------------------------------------------------------------------------------
# NAMLModule against the golden fixtures minted from the reference modules
# ------------------------------------------------------------------------------------------------
def load_naml_golden(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    V, E, F_, W, Q, CE, C, B, max_hist, seed, L, LA = [int(x) for x in g["meta"]]
    if any(k.startswith("param/") for k in g):
        params = {k[len("param/"):]: torch.from_numpy(g[k]) for k in g if k.startswith("param/")}
    else:
        params = make_naml_params(V, E, F_, W, Q, CE, C, seed=seed)
        chk = np.array([float(v.double().sum()) for v in params.values()])
        assert np.allclose(chk, g["param_checksum"], rtol=1e-9)
    batch = {"batch_hist": torch.from_numpy(g["batch_hist"]), "batch_cand": torch.from_numpy(g["batch_cand"]),
             "labels": torch.from_numpy(g["labels"]), "x_hist": {}, "x_cand": {}}
    for side in ("hist", "cand"):
        for attr in ("title", "abstract", "category"):
            batch["x_" + side][attr] = torch.from_numpy(g[f"{side}_{attr}"])
    batch["user_idx"] = torch.arange(B)
    return g, params, batch, dict(V=V, E=E, F=F_, W=W, Q=Q, CE=CE, C=C, B=B)


def make_naml_module(params, d, late_fusion=False, p=0.2):
    from newsreclib_b200.models.general_rec.naml_module import NAMLModule
    m = NAMLModule(
        dataset_attributes=["title", "abstract", "category", "subcategory"],
        attributes2encode=["title", "abstract", "category"], outputs=OUTPUTS, dual_loss_training=False,
        dual_loss_coef=None, loss="cross_entropy_loss", late_fusion=late_fusion, temperature=None, use_plm=False,
        pretrained_embeddings_path=None, plm_model=None, frozen_layers=None, text_embed_dim=d["E"], num_heads=15,
        num_filters=d["F"], window_size=d["W"], query_dim=d["Q"], categ_embed_dim=d["CE"], dropout_probability=p,
        top_k_list=[5, 10], num_categ_classes=d["C"] - 1, num_sent_classes=3, save_recs=False, recs_fpath=None,
        optimizer=functools.partial(torch.optim.Adam, lr=1e-4), scheduler=None,
        pretrained_embeddings=params[T_ + "embedding_layer.weight"])
    full = dict(params)
    for k in list(params):
        if ".text_encoders.title." in k:
            full[k.replace(".title.", ".abstract.")] = params[k]
    if late_fusion:
        full = {k: v for k, v in full.items() if not k.startswith("user_encoder.")}
    res = m.load_state_dict(full, strict=True)  # the reference's key names, both aliased prefixes
    assert not res.missing_keys and not res.unexpected_keys
    return m


def naml_dev_batch(batch, dev="cuda"):
    b = {k: (v.to(dev) if torch.is_tensor(v) else {kk: vv.to(dev) for kk, vv in v.items()} if isinstance(v, dict) else v)
         for k, v in batch.items()}
    return b


@pytest.mark.parametrize("name", ["naml_tiny", "naml_mind"])
def test_naml_module_matches_reference_golden(name):
    g, params, batch, d = load_naml_golden(name)
    m = make_naml_module(params, d).cuda().eval()
    b = naml_dev_batch(batch)
    scores = m(b)
    assert rel_err(scores, g["scores"]) <= 1e-4
    cnt = torch.bincount(batch["batch_cand"], minlength=d["B"])
    for i in range(d["B"]):
        assert torch.all(scores[i, cnt[i]:] == 0)  # padded candidate slots score exactly 0
    out = m.model_step(b)
    assert len(out) == 11
    assert rel_err(out[0], g["loss"]) <= 1e-4
    out[0].backward()

    def f(leaves, dt):
        bb = dict(batch); bb["labels"] = batch["labels"].to(dt)
        return O.nrms_loss(bb, O.naml_forward(batch, leaves, d["W"]))
    _, tol = oracle_grads(f, params)
    grads = {k.replace(".abstract.", ".title."): v.grad for k, v in m.named_parameters()}
    for k, v in g.items():
        if k.startswith("grad/"):
            assert rel_err(grads[k[5:]], v) <= tol[k[5:]], k
        elif k.startswith("gradsample/"):
            assert rel_err(grads[k[11:]].reshape(-1)[::7], v) <= tol[k[11:]], k
    assert float(grads[T_ + "embedding_layer.weight"][0].abs().max()) == 0.0
    assert float(grads[C_ + "embedding_layer.weight"][0].abs().max()) == 0.0


def test_naml_module_late_fusion_and_training_steps():
    V = 800
    d = dict(V=V, E=300, F=400, W=3, Q=200, CE=100, C=19, B=6)
    params = make_naml_params(V, seed=8)
    batch = make_batch(6, V, hist="ragged", max_hist=7, cand="train", seed=8, abstract_len=50)
    m = make_naml_module(params, d, late_fusion=True).cuda().eval()
    b = naml_dev_batch(batch)
    ref = O.naml_forward(batch, params, 3, late_fusion=True)
    assert rel_err(m(b), ref) <= 1e-4
    m = make_naml_module(params, d).cuda().train()
    opt = m.configure_optimizers()["optimizer"]
    losses = []
    for step in range(6):
        opt.zero_grad()
        loss = m.training_step(b, step)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


# ------------------------------------------------------------------------------------------------
# PLM head (a14)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["plm_head_d48", "plm_head_d64"])
def test_plm_head_matches_reference_golden(name):
    from newsreclib_b200 import ops
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    hidden, heads, Q, N, T = [int(x) for x in g["meta"]]
    p = {k[len("param/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param/")}
    dev = {k: p["multihead_attention." + k if not k.startswith("additive") else k].cuda().requires_grad_(True)
           for k in ["in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias",
                     "additive_attention.linear.weight", "additive_attention.linear.bias", "additive_attention.query"]}
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    out = ops.PlmHeadFn.apply(x, *dev.values(), heads, 0, 0.0, False, 0, ops.PREC_BF16X3)
    assert rel_err(out, g["out"]) <= 1e-4
    (out * torch.from_numpy(g["w"]).cuda()).sum().backward()
    x64 = torch.from_numpy(g["x"])

    def f(leaves, dt):
        return (O.plm_head(leaves["x"], leaves, heads) * torch.from_numpy(g["w"]).to(dt)).sum()
    _, tol = oracle_grads(f, {"x": x64, **p})
    assert rel_err(x.grad, g["dx"]) <= tol["x"]
    for k, v in dev.items():
        full = k if k.startswith("additive") else "multihead_attention." + k
        assert rel_err(v.grad, g["grad/" + full]) <= tol[full], k


def test_plm_head_dropout_and_token_axis():
    from newsreclib_b200 import ops
    g = dict(np.load(os.path.join(GOLD, "plm_head_d48.npz")))
    hidden, heads, Q, N, T = [int(x) for x in g["meta"]]
    p = {k[len("param/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param/")}
    order = ["multihead_attention.in_proj_weight", "multihead_attention.in_proj_bias",
             "multihead_attention.out_proj.weight", "multihead_attention.out_proj.bias",
             "additive_attention.linear.weight", "additive_attention.linear.bias", "additive_attention.query"]
    dev = [p[k].cuda() for k in order]
    x = torch.from_numpy(g["x"])
    pd, seed = 0.2, 99
    out = ops.PlmHeadFn.apply(x.cuda(), *dev, heads, 0, pd, True, seed, ops.PREC_BF16X3)
    k1 = ops.dropout_mask(N * T * hidden, seed, 0, pd, "cuda").cpu().float().reshape(N, T, hidden)
    k2 = ops.dropout_mask(N * T * hidden, seed, 1, pd, "cuda").cpu().float().reshape(N, T, hidden)
    assert rel_err(out, O.plm_head(x, p, heads, k1, k2, pd)) <= 1e-4
    # attention along the tokens instead of across the news (attention_axis=1)
    out_tok = ops.PlmHeadFn.apply(x.cuda(), *dev, heads, 1, 0.0, False, 0, ops.PREC_BF16X3)
    y = O.mha_seq_first(x.permute(1, 0, 2), *[p[k] for k in order[:4]], heads).permute(1, 0, 2)
    ref_tok = O.additive_attention(y, p[order[4]], p[order[5]], p[order[6]])
    assert rel_err(out_tok, ref_tok) <= 1e-4


@pytest.mark.parametrize("hidden,heads,N,T,axis", [(768, 16, 400, 6, 0), (128, 2, 200, 5, 0), (768, 16, 9, 96, 1),
                                                   (768, 16, 130, 3, 0)])
def test_plm_head_long_sequences_vs_oracle(hidden, heads, N, T, axis):
    """The PLM head at the sequence lengths the reference's batch-axis attention produces (S = N = the 400 clicked news
    of 8 impressions; head dims 48 and 64; S not a multiple of the 64-key / 128-query blocks) and along 96 tokens:
    the flash-style tensor-core attention kernels (nrl_attn_flash.cuh), forward and backward, against the oracle."""
    from newsreclib_b200 import ops
    g = torch.Generator().manual_seed(N + T)
    Q = 40
    order = ["multihead_attention.in_proj_weight", "multihead_attention.in_proj_bias",
             "multihead_attention.out_proj.weight", "multihead_attention.out_proj.bias",
             "additive_attention.linear.weight", "additive_attention.linear.bias", "additive_attention.query"]
    shapes = [(3 * hidden, hidden), (3 * hidden,), (hidden, hidden), (hidden,), (Q, hidden), (Q,), (Q,)]
    p = {k: torch.randn(*sh, generator=g) * (0.08 if len(sh) == 2 else 0.1) for k, sh in zip(order, shapes)}
    x = torch.randn(N, T, hidden, generator=g)
    wgt = torch.randn(N, hidden, generator=g)
    dev = [p[k].cuda().requires_grad_(True) for k in order]
    xd = x.cuda().requires_grad_(True)
    out = ops.PlmHeadFn.apply(xd, *dev, heads, axis, 0.0, False, 0, ops.PREC_BF16X3)
    (out * wgt.cuda()).sum().backward()

    def ref(leaves, dt):
        if axis == 0:
            o = O.plm_head(leaves["x"], leaves, heads)
        else:
            y = O.mha_seq_first(leaves["x"].permute(1, 0, 2), *[leaves[k] for k in order[:4]], heads).permute(1, 0, 2)
            o = O.additive_attention(y, leaves[order[4]], leaves[order[5]], leaves[order[6]])
        ref.out = o.detach() if dt == torch.float32 else ref.out
        return (o * wgt.to(dt)).sum()
    ref.out = None
    rg, tol = oracle_grads(ref, {"x": x, **p})
    e = rel_err(out, ref.out)
    errs = {"x": rel_err(xd.grad, rg["x"]) / tol["x"]}
    errs.update({k: rel_err(v.grad, rg[k]) / tol[k] for k, v in zip(order, dev)})
    worst = max(errs, key=errs.get)
    print(f"[plm head] E={hidden} heads={heads} N={N} T={T} axis={axis}: out rel {e:.2e} (tol 1e-4); worst gradient error / "
          f"tolerance {errs[worst]:.2f} ({worst}, tol {tol[worst]:.1e})")
    assert e <= 1e-4
    assert errs[worst] <= 1.0, (worst, errs[worst])


def test_nrms_plm_module_forward():
    """NRMSModule(use_plm=True, transformer_impl="hf") around a tiny random RoBERTa with head dim 48 (outside the sm_100a
    transformer's coverage): HF transformer on torch + the sm_100a head.  The native transformer: tests/test_gpu_tfm.py."""
    from transformers import RobertaConfig, RobertaModel
    from newsreclib_b200.models.general_rec.nrms_module import NRMSModule
    torch.manual_seed(0)
    hidden, T = 96, 12
    cfg = RobertaConfig(vocab_size=120, hidden_size=hidden, num_hidden_layers=2, num_attention_heads=2,
                        intermediate_size=192, max_position_embeddings=T + 4, pad_token_id=1)
    tf = RobertaModel(cfg).eval()
    m = NRMSModule(
        dataset_attributes=["title"], attributes2encode=["title"], outputs=OUTPUTS, dual_loss_training=False,
        dual_loss_coef=None, loss="cross_entropy_loss", late_fusion=False, temperature=None, use_plm=True,
        pretrained_embeddings_path=None, plm_model=tf, frozen_layers=[0], embed_dim=hidden, num_heads=2,
        query_dim=40, dropout_probability=0.2, top_k_list=[5], num_categ_classes=18, num_sent_classes=3,
        save_recs=False, recs_fpath=None, optimizer=None, scheduler=None, transformer_impl="hf")  # head dim 48
    frozen = [n for n, q in m.named_parameters() if not q.requires_grad]
    assert frozen and all("layer.0." in n for n in frozen)
    B = 4
    hcnt, ccnt = [3, 5, 2, 4], [5, 5, 5, 5]
    g = torch.Generator().manual_seed(1)

    def text(n):
        lens = torch.randint(4, T + 1, (n,), generator=g); lens[0] = T
        att = (torch.arange(T)[None, :] < lens[:, None]).long()
        ids = torch.where(att.bool(), torch.randint(3, 120, (n, T), generator=g), torch.ones(n, T, dtype=torch.long))
        return {"input_ids": ids, "attention_mask": att}
    batch = {"x_hist": {"title": text(sum(hcnt))}, "x_cand": {"title": text(sum(ccnt))},
             "batch_hist": torch.repeat_interleave(torch.arange(B), torch.tensor(hcnt)),
             "batch_cand": torch.repeat_interleave(torch.arange(B), torch.tensor(ccnt)),
             "labels": torch.tensor([1., 0, 0, 0, 0] * B), "user_idx": torch.arange(B)}
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        te = "news_encoder.text_encoders.title."
        hist = O.plm_head(tf(**batch["x_hist"]["title"])[0], sub(sd, te), 2)
        cand = O.plm_head(tf(**batch["x_cand"]["title"])[0], sub(sd, te), 2)
        u = O.nrms_user_encoder(O.to_dense_batch(hist, batch["batch_hist"])[0], sub(sd, "user_encoder."), 2)
        ref = O.dot_product(u, O.to_dense_batch(cand, batch["batch_cand"])[0])
    mc = m.cuda().eval()
    b = {k: (v.cuda() if torch.is_tensor(v) else {"title": {kk: vv.cuda() for kk, vv in v["title"].items()}})
         for k, v in batch.items()}
    scores = mc(b)
    assert rel_err(scores, ref) <= 2e-4  # includes the fp32 GPU-vs-CPU noise of the HF transformer
    loss = mc.model_step(b)[0]
    loss.backward()
    got = [n for n, q in mc.named_parameters() if q.grad is not None]
    assert any("plm_model" in n for n in got) and any("multihead_attention" in n for n in got)


def test_module_trainer_matches_torch_adam_on_naml():
    """ModuleTrainer (flat buffers + nrl_adam_step) follows torch.optim.Adam on the same NAML module."""
    from newsreclib_b200.trainer import ModuleTrainer
    V = 600
    d = dict(V=V, E=300, F=400, W=3, Q=200, CE=100, C=19, B=6)
    params = make_naml_params(V, seed=12)
    batch = make_batch(6, V, hist="ragged", max_hist=6, cand="train", seed=12, abstract_len=50)
    b = naml_dev_batch(batch)
    ma = make_naml_module(params, d, p=0.0).cuda()
    mb = make_naml_module(params, d, p=0.0).cuda()
    tr = ModuleTrainer(ma, lr=1e-3)
    opt = torch.optim.Adam(mb.parameters(), lr=1e-3)
    for step in range(3):
        la = tr.train_step(b)
        opt.zero_grad()
        lb = mb.training_step(b, step)
        lb.backward()
        opt.step()
        assert rel_err(la, lb.detach()) <= 1e-5
    # Adam's step is lr * m / (sqrt(v) + eps) ~ +-lr wherever |g| >> eps, so elements whose gradient is ~0
    # (sign decided by the last bits of a split-K reduction) may legitimately differ by O(lr) between two
    # runs; everything else agrees to fp32 noise
    sa, sb = ma.state_dict(), mb.state_dict()
    for k in sa:
        diff = (sa[k] - sb[k]).abs()
        assert float(diff.max()) <= 3 * 2e-3 + 1e-6, k                      # never more than the 3 steps' worth
        assert float((diff > 1e-5).float().mean()) <= 0.02, k                # and only on a sliver of elements
