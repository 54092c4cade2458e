"""Pin the oracle against the reference's own modules and mint golden fixtures.

Runs ONLY in the build container (it imports the unmodified reference from
``/root/reference``; that path does not exist on the GPU box).  For every config it

1. builds the reference ``MHSAAddAtt`` / ``NewsEncoder`` / NRMS ``UserEncoder`` /
   ``DotProduct`` (+ NAML ``CNNAddAtt`` / ``LinearEncoder`` / NAML ``UserEncoder``) and loads
   seeded weights into them,
2. runs the reference forward + ``CrossEntropyLoss`` + backward in eval mode,
3. runs ``oracle/nrms_oracle.py`` on the same inputs (autograd over the restatement),
4. asserts agreement (fp32 noise floor) and writes inputs/outputs/gradients to
   ``tests/golden/*.npz``.

Usage:  python oracle/make_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from newsreclib.models.components.encoders.news.news import NewsEncoder  # noqa: E402
from newsreclib.models.components.encoders.news.text import CNNAddAtt, MHSAAddAtt  # noqa: E402
from newsreclib.models.components.encoders.news.category import LinearEncoder  # noqa: E402
from newsreclib.models.components.encoders.user.naml import UserEncoder as NAMLUser  # noqa: E402
from newsreclib.models.components.encoders.user.nrms import UserEncoder as NRMSUser  # noqa: E402
from newsreclib.models.components.layers.click_predictor import DotProduct  # noqa: E402

from newsreclib_b200.synthetic import make_batch, make_nrms_params  # noqa: E402
from oracle import nrms_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def reference_nrms(params, batch, E, H, Q):
    """The reference modules glued exactly as ``nrms_module.py:230-255`` + ``:277,288``."""
    table = params["news_encoder.text_encoders.title.embedding_layer.weight"]
    text = MHSAAddAtt(table.clone(), E, H, Q, 0.2)
    news = NewsEncoder(["title"], ["title"], False, text, None, None, False, None, None, None, None)
    user = NRMSUser(E, H, Q)
    mod = torch.nn.Module()
    mod.news_encoder, mod.user_encoder = news, user
    missing = mod.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    mod.eval()
    hist = news(batch["x_hist"])
    hist_agg, _ = O.to_dense_batch(hist, batch["batch_hist"])
    cand = news(batch["x_cand"])
    cand_agg, _ = O.to_dense_batch(cand, batch["batch_cand"])
    u = user(hist_agg)
    scores = DotProduct()(u.unsqueeze(1), cand_agg.permute(0, 2, 1))
    y, _ = O.to_dense_batch(batch["labels"], batch["batch_cand"])
    loss = torch.nn.CrossEntropyLoss()(scores, y)
    loss.backward()
    grads = {k: v.grad.detach().clone() for k, v in mod.named_parameters()}
    return dict(hist=hist.detach(), cand=cand.detach(), user=u.detach(),
                scores=scores.detach(), loss=loss.detach(), grads=grads)


def oracle_nrms(params, batch, H):
    ps = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    title, user = O.split_params(ps)
    hist = O.mhsa_add_att(batch["x_hist"]["title"], title, H)
    hist_agg, _ = O.to_dense_batch(hist, batch["batch_hist"])
    cand = O.mhsa_add_att(batch["x_cand"]["title"], title, H)
    cand_agg, _ = O.to_dense_batch(cand, batch["batch_cand"])
    u = O.nrms_user_encoder(hist_agg, user, H)
    scores = O.dot_product(u, cand_agg)
    loss = O.nrms_loss(batch, scores)
    loss.backward()
    grads = {k: v.grad.detach().clone() for k, v in ps.items()}
    # padding_idx=0: the reference zeroes the gradient of table row 0 (text.py:215-217)
    grads["news_encoder.text_encoders.title.embedding_layer.weight"][0] = 0
    return dict(hist=hist.detach(), cand=cand.detach(), user=u.detach(),
                scores=scores.detach(), loss=loss.detach(), grads=grads)


def nrms_case(name, E, H, Q, V, B, max_hist, hist, cand, seed, store_params, max_title_len=30):
    params = make_nrms_params(V, E, H, Q, seed=seed)
    batch = make_batch(B, V, hist=hist, max_hist=max_hist, cand=cand, seed=seed,
                       max_title_len=max_title_len)
    if cand == "eval":  # keep the fixture small
        pass
    ref = reference_nrms(params, batch, E, H, Q)
    orc = oracle_nrms(params, batch, H)
    worst = 0.0
    for k in ("hist", "cand", "user", "scores", "loss"):
        r = rel(orc[k], ref[k]); worst = max(worst, r)
        assert r < 2e-5, (name, k, r)
    for k, g in ref["grads"].items():
        r = rel(orc["grads"][k], g); worst = max(worst, r)
        assert r < 2e-4, (name, "grad", k, r)
    assert float(ref["grads"]["news_encoder.text_encoders.title.embedding_layer.weight"][0].abs().max()) == 0.0
    out = {
        "meta": np.array([E, H, Q, V, B, max_hist, seed, max_title_len], dtype=np.int64),
        "hist_mode": np.array(hist), "cand_mode": np.array(cand),
        "oracle_vs_reference_maxrel": np.array(worst),
        "hist_title": batch["x_hist"]["title"].numpy(), "cand_title": batch["x_cand"]["title"].numpy(),
        "batch_hist": batch["batch_hist"].numpy(), "batch_cand": batch["batch_cand"].numpy(),
        "labels": batch["labels"].numpy(),
        "hist_vec": ref["hist"].numpy(), "cand_vec": ref["cand"].numpy(), "user_vec": ref["user"].numpy(),
        "scores": ref["scores"].numpy(), "loss": ref["loss"].numpy(),
        "param_checksum": np.array([float(v.double().sum()) for v in params.values()]),
    }
    for k, g in ref["grads"].items():
        g = g.numpy()
        if store_params or g.size <= 4096 or k.endswith("embedding_layer.weight"):
            out["grad/" + k] = g
        else:  # strided sample + moments of the big projection-weight gradients
            out["gradsample/" + k] = g.reshape(-1)[::7].copy()
            out["gradsum/" + k] = np.array([g.astype(np.float64).sum(), np.abs(g).astype(np.float64).sum()])
    if store_params:
        for k, v in params.items():
            out["param/" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: oracle vs reference max rel {worst:.2e}; loss {float(ref['loss']):.6f}; "
          f"N_h={batch['batch_hist'].numel()} N_c={batch['batch_cand'].numel()}")


def coupling_case():
    """Pin the batch-axis attention quirk (``encoders/user/nrms.py:34-36``): perturbing
    user 1's history changes user 0's vector."""
    torch.manual_seed(7)
    E, H, Q = 60, 3, 40
    user = NRMSUser(E, H, Q).eval()
    h = torch.randn(3, 5, E)
    h2 = h.clone(); h2[1] += 0.5
    with torch.no_grad():
        u, u2 = user(h), user(h2)
        p = {k: v for k, v in user.state_dict().items()}
        o, o2 = O.nrms_user_encoder(h, p, H), O.nrms_user_encoder(h2, p, H)
    assert rel(o, u) < 1e-5 and rel(o2, u2) < 1e-5
    delta = float((u2[0] - u[0]).abs().max())
    assert delta > 1e-3
    np.savez_compressed(os.path.join(GOLD, "user_coupling.npz"), h=h.numpy(), h2=h2.numpy(),
                        u=u.numpy(), u2=u2.numpy(), meta=np.array([E, H, Q]),
                        **{"param/" + k: v.numpy() for k, v in p.items()})
    print(f"user_coupling: delta on user 0 = {delta:.4f}")


def naml_case():
    """NAML news encoder views + user encoder (``naml_module.py:261-286``)."""
    torch.manual_seed(11)
    V, E, F_, W, Q, C, CE = 80, 300, 400, 3, 200, 19, 100
    table = torch.randn(V + 1, E)
    text = CNNAddAtt(table.clone(), E, F_, W, Q, 0.2)
    categ = LinearEncoder(pretrained_embeddings=None, from_pretrained=False, freeze_pretrained_emb=False,
                          num_categories=C, embed_dim=CE, use_dropout=True, dropout_probability=0.2,
                          linear_transform=True, output_dim=F_)
    news = NewsEncoder(["title", "abstract", "category"], ["title", "abstract", "category"], False,
                       text, categ, None, True, "add_att", F_, Q, None).eval()
    user = NAMLUser(F_, Q).eval()
    rng = np.random.default_rng(5)
    from newsreclib_b200.synthetic import make_titles
    n = 7
    x = {"title": torch.from_numpy(make_titles(rng, n, V, 30)),
         "abstract": torch.from_numpy(make_titles(rng, n, V, 50, mean_len=40.0, min_len=0)),
         "category": torch.from_numpy(rng.integers(1, C, n).astype(np.int64))}
    with torch.no_grad():
        vec = news(x)
        tp = {k[len("text_encoders.title."):]: v for k, v in news.state_dict().items()
              if k.startswith("text_encoders.title.")}
        cp = {k[len("category_encoders.category."):]: v for k, v in news.state_dict().items()
              if k.startswith("category_encoders.category.")}
        lp = {k[len("combine_layer."):]: v for k, v in news.state_dict().items()
              if k.startswith("combine_layer.")}
        views = {name: O.cnn_add_att(x[name], tp, W) for name in ("title", "abstract")}
        views["category"] = O.linear_category_encoder(x["category"], cp)
        # same ModuleDict order as the reference instance (set-iteration order, news.py:68-77)
        order = [k for k in news.text_encoders.keys()] + [k for k in news.category_encoders.keys()]
        stacked = torch.stack([views[k] for k in order], dim=1)
        ovec = O.additive_attention(stacked, lp["linear.weight"], lp["linear.bias"], lp["query"])
        r = rel(ovec, vec)
        assert r < 2e-5, r
        h = torch.randn(2, 4, F_)
        uref = user(h)
        up = {k: v for k, v in user.state_dict().items()}
        r2 = rel(O.naml_user_encoder(h, up), uref)
        assert r2 < 2e-5, r2
    np.savez_compressed(
        os.path.join(GOLD, "naml_news.npz"), title=x["title"].numpy(), abstract=x["abstract"].numpy(),
        category=x["category"].numpy(), news_vec=vec.numpy(), user_in=h.numpy(), user_vec=uref.numpy(),
        meta=np.array([V, E, F_, W, Q, C, CE]),
        **{"news/" + k: v.numpy() for k, v in news.state_dict().items() if not k.startswith("text_encoders.abstract.")},
        **{"user/" + k: v.numpy() for k, v in up.items()})
    print(f"naml_news: oracle vs reference rel {r:.2e} (news), {r2:.2e} (user)")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(4)
    # tiny dims: everything stored (params, all grads)
    nrms_case("nrms_tiny", E=60, H=3, Q=40, V=50, B=3, max_hist=5, hist="ragged", cand="train",
              seed=3, store_params=True, max_title_len=12)
    # the real NRMS dims (configs/model/nrms.yaml:19-21), params regenerated from the seed
    nrms_case("nrms_mind", E=300, H=15, Q=200, V=300, B=4, max_hist=8, hist="ragged", cand="train",
              seed=1234, store_params=False)
    # config 1 of BASELINE.json: batch_size=8 plumbing shape (fixed history kept short)
    nrms_case("nrms_b8", E=300, H=15, Q=200, V=500, B=8, max_hist=10, hist="fixed", cand="train",
              seed=42, store_params=False)
    coupling_case()
    naml_case()
