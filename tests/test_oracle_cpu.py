"""CPU suite (-m "not gpu"): the oracle against the committed golden fixtures (minted from the
reference's own modules by oracle/make_golden.py), host logic, and the C-ABI export check."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from helpers import GOLD, TITLE, USER, grad_tolerances, load_golden, oracle_run, rel_err
from oracle import nrms_oracle as O


@pytest.mark.parametrize("name", ["nrms_tiny", "nrms_mind", "nrms_b8"])
def test_oracle_matches_reference_golden(name):
    g, params, batch, d = load_golden(name)
    assert float(g["oracle_vs_reference_maxrel"]) < 2e-4
    scores, loss, grads = oracle_run(params, batch, d["H"])
    # fp32 noise floor between two CPU evaluations of the same math
    assert rel_err(scores, g["scores"]) < 1e-5
    assert rel_err(loss, g["loss"]) < 1e-6
    title, _ = O.split_params(params)
    assert rel_err(O.mhsa_add_att(batch["x_hist"]["title"], title, d["H"]), g["hist_vec"]) < 1e-5
    # per-tensor bar = max(2e-4, 4 x the fp32 oracle's own error against an fp64 run): the
    # additive-attention bias gradient is a sum that cancels to ~1e-5 of its terms, so two fp32
    # evaluations on different hosts (BLAS thread counts) differ by more than 2e-4 there
    tol = grad_tolerances(params, batch, d["H"], 2e-4, ref_grads=grads)
    for k, v in g.items():
        if k.startswith("grad/"):
            assert rel_err(grads[k[5:]], v) < tol[k[5:]], k
        elif k.startswith("gradsample/"):
            assert rel_err(grads[k[11:]].reshape(-1)[::7], v) < tol[k[11:]], k
    # padded candidate slots score exactly 0, embedding row 0 has exactly zero gradient
    B = d["B"]
    cnt = torch.bincount(batch["batch_cand"], minlength=B)
    for b in range(B):
        assert torch.all(scores[b, cnt[b]:] == 0)
    assert float(grads[TITLE + "embedding_layer.weight"][0].abs().max()) == 0.0


def test_embedding_gather_bit_exact_including_row0():
    table = torch.randn(17, 12)
    ids = torch.tensor([[0, 3, 16], [0, 0, 5]])
    out = O.embedding_gather(table, ids)
    assert torch.equal(out[0, 0], table[0]) and torch.equal(out[1, 2], table[5])
    assert out[0, 0].abs().sum() > 0  # row 0 is a real row, not zeros


def test_to_dense_batch_semantics():
    x = torch.arange(12.0).reshape(6, 2)
    batch = torch.tensor([0, 0, 0, 2, 2, 3])  # segment 1 is empty
    dense, mask = O.to_dense_batch(x, batch)
    assert dense.shape == (4, 3, 2) and mask.shape == (4, 3)
    assert torch.equal(dense[0], x[:3]) and torch.equal(dense[2, :2], x[3:5])
    assert torch.all(dense[1] == 0) and not mask[1].any()
    assert mask.sum() == 6


def test_user_coupling_quirk():
    g = dict(np.load(os.path.join(GOLD, "user_coupling.npz")))
    E, H, Q = [int(x) for x in g["meta"]]
    p = {k[len("param/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param/")}
    u = O.nrms_user_encoder(torch.from_numpy(g["h"]), p, H)
    u2 = O.nrms_user_encoder(torch.from_numpy(g["h2"]), p, H)
    assert rel_err(u, g["u"]) < 1e-5 and rel_err(u2, g["u2"]) < 1e-5
    # perturbing impression 1 changes impression 0 (batch_first=False quirk, user/nrms.py:34-36)
    assert float((u2[0] - u[0]).abs().max()) > 1e-3


def test_naml_golden():
    g = dict(np.load(os.path.join(GOLD, "naml_news.npz")))
    V, E, F_, W, Q, C_, CE = [int(x) for x in g["meta"]]
    news = {k[len("news/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("news/")}
    tp = {k[len("text_encoders.title."):]: v for k, v in news.items() if k.startswith("text_encoders.title.")}
    cp = {k[len("category_encoders.category."):]: v for k, v in news.items() if k.startswith("category_encoders.category.")}
    lp = {k[len("combine_layer."):]: v for k, v in news.items() if k.startswith("combine_layer.")}
    views = [O.cnn_add_att(torch.from_numpy(g["title"]), tp, W), O.cnn_add_att(torch.from_numpy(g["abstract"]), tp, W),
             O.linear_category_encoder(torch.from_numpy(g["category"]), cp)]
    vec = O.additive_attention(torch.stack(views, 1), lp["linear.weight"], lp["linear.bias"], lp["query"])
    assert rel_err(vec, g["news_vec"]) < 2e-5  # view pooling is permutation invariant
    up = {k[len("user/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("user/")}
    assert rel_err(O.naml_user_encoder(torch.from_numpy(g["user_in"]), up), g["user_vec"]) < 2e-5


@pytest.mark.parametrize("name", ["naml_tiny", "naml_mind"])
def test_naml_step_oracle_matches_reference_golden(name):
    """Whole NAML step (naml_module.py:261-286 + CE + backward) minted from the reference modules."""
    from newsreclib_b200.synthetic import make_naml_params
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    V, E, F_, W, Q, CE, C, B, max_hist, seed, L, LA = [int(x) for x in g["meta"]]
    assert float(g["oracle_vs_reference_maxrel"]) < 5e-4
    if any(k.startswith("param/") for k in g):
        params = {k[len("param/"):]: torch.from_numpy(g[k]) for k in g if k.startswith("param/")}
    else:
        params = make_naml_params(V, E, F_, W, Q, CE, C, seed=seed)
        chk = np.array([float(v.double().sum()) for v in params.values()])
        assert np.allclose(chk, g["param_checksum"], rtol=1e-9), "seeded parameter generator drifted"
    batch = {"batch_hist": torch.from_numpy(g["batch_hist"]), "batch_cand": torch.from_numpy(g["batch_cand"]),
             "labels": torch.from_numpy(g["labels"]), "x_hist": {}, "x_cand": {}}
    for side in ("hist", "cand"):
        for attr in ("title", "abstract", "category"):
            batch["x_" + side][attr] = torch.from_numpy(g[f"{side}_{attr}"])
    ps = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    scores = O.naml_forward(batch, ps, W)
    loss = O.nrms_loss(batch, scores)
    assert rel_err(scores.detach(), g["scores"]) < 1e-5 and rel_err(loss.detach(), g["loss"]) < 1e-6
    loss.backward()
    for k, v in g.items():
        if k.startswith("grad/") and not k.endswith("embedding_layer.weight"):
            assert rel_err(ps[k[5:]].grad, v) < 2e-3, k
    # multi-view news vectors and the additive user vector
    news = {k[len("news_encoder."):]: v for k, v in params.items() if k.startswith("news_encoder.")}
    assert rel_err(O.naml_news_encoder(batch["x_hist"], news, W), g["hist_vec"]) < 1e-5


@pytest.mark.parametrize("name", ["plm_head_d48", "plm_head_d64"])
def test_plm_head_oracle_matches_reference_golden(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    hidden, heads, Q, N, T = [int(x) for x in g["meta"]]
    p = {k[len("param/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param/")}
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    out = O.plm_head(x, p, heads)
    assert rel_err(out.detach(), g["out"]) < 1e-5
    (out * torch.from_numpy(g["w"])).sum().backward()
    assert rel_err(x.grad, g["dx"]) < 2e-4
    # the quirk (text.py:96): attention runs across the N news -> perturbing news 1 changes news 0
    x2 = torch.from_numpy(g["x"]).clone(); x2[1] += 0.5
    assert float((O.plm_head(x2, p, heads)[0] - out.detach()[0]).abs().max()) > 1e-4


def test_ce_soft_matches_torch_and_multi_positive():
    s = torch.randn(5, 7)
    y = torch.zeros(5, 7); y[:, 1] = 1; y[2, 4] = 1  # a multi-positive row (sum y > 1)
    assert rel_err(O.ce_soft(s, y), torch.nn.CrossEntropyLoss()(s, y)) < 1e-6


def test_adam_restatement_matches_torch():
    torch.manual_seed(0)
    p0 = torch.randn(1000)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-4)
    p, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    for step in range(1, 4):
        g = torch.randn(1000) * (step == 2 and 0.0 or 1.0)  # a zero-gradient step still moves p
        p_ref.grad = g.clone(); opt.step()
        O.adam_step(p, g, m, v, step)
    assert rel_err(p, p_ref.detach()) < 1e-6


def test_synthetic_batch_layout():
    from newsreclib_b200.synthetic import make_batch
    b = make_batch(8, 1000, hist="ragged", cand="eval", seed=5)
    assert b["x_hist"]["title"].shape[1] == 30 and b["x_hist"]["title"].dtype == torch.int64
    assert torch.all(b["batch_hist"][1:] >= b["batch_hist"][:-1])  # sorted segment ids
    assert int(b["batch_hist"].max()) == 7 and b["labels"].dtype == torch.float32
    assert int(b["x_hist"]["title"].max()) <= 1000 and int(b["x_hist"]["title"].min()) == 0
    cnt = torch.bincount(b["batch_cand"])
    assert cnt.min() >= 2 and cnt.max() <= 300


def test_cabi_exports_every_declared_symbol():
    """The shared library loads and exports every function include/nrl.h declares."""
    from newsreclib_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "nrl.h")).read()
    declared = set(re.findall(r"\b(nrl_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/nrl.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib.nrl_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.nrl_version()


def test_product_path_refuses_cpu_tensors():
    from newsreclib_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.gemm_test(torch.zeros(4, 16), torch.zeros(4, 16), False)
    # the optimizer and the loss / metric wrappers have no host path either
    from newsreclib_b200.optim import Adam
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        Adam([p], lr=1e-3).step()
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.SupConFn.apply(torch.zeros(2, 3), torch.zeros(6), torch.tensor([0, 3, 6], dtype=torch.int32), None)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.rank_metrics(torch.zeros(6), torch.zeros(6), torch.tensor([3, 3]), [5])


def test_hydra_model_configs_match_module_constructors():
    """configs/model/{nrms,naml}_b200.yaml carry exactly the constructor kwargs of the drop-in modules
    (the reference's configs/model/{nrms,naml}.yaml keys) and point _target_ at them."""
    import importlib
    import inspect
    import yaml
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name in ("nrms_b200", "naml_b200"):
        cfg = yaml.safe_load(open(os.path.join(root, "configs", "model", name + ".yaml")))
        mod, cls = cfg.pop("_target_").rsplit(".", 1)
        klass = getattr(importlib.import_module(mod), cls)
        params = inspect.signature(klass.__init__).parameters
        required = {k for k, v in params.items() if k != "self" and v.default is inspect.Parameter.empty}
        assert required == set(cfg), (name, required ^ set(cfg))


@pytest.mark.parametrize("split", ["test", "train"])
def test_collate_oracle_matches_reference_collate_golden(split):
    """collate_ref.npz holds the output of the reference's OWN DatasetCollate.__call__ (rec_dataset.py:148-293) on
    impressions produced by its own Dataset classes (test split; train split with 4:1 negative sampling under a fixed
    numpy seed): the restatement must reproduce every tensor bit for bit, dtype included."""
    from helpers import load_collate_golden
    from oracle import collate_oracle as CO
    news, splits, (lt, la) = load_collate_golden()
    samples, ref = splits[split]
    got = CO.collate(news, samples, lt, la)
    assert set(got) == set(ref)
    for k, v in ref.items():
        if isinstance(v, dict):
            assert set(got[k]) == set(v)
            for c, t in v.items():
                assert got[k][c].dtype == t.dtype and torch.equal(got[k][c], t), (k, c)
        else:
            assert got[k].dtype == v.dtype and torch.equal(got[k], v), k
    assert ref["x_hist"]["title"].shape[1] == lt and ref["x_hist"]["abstract"].shape[1] == la
    assert int(torch.bincount(ref["batch_hist"]).max()) <= 50          # max_history_len applied by the Dataset
    if split == "train":                                               # 1 positive : 4 negatives per positive
        B = int(ref["batch_cand"].max()) + 1
        for b in range(B):
            y = ref["labels"][ref["batch_cand"] == b]
            assert int((y == 0).sum()) == 4 * int((y == 1).sum())


def test_device_table_padding_matches_reference_collate_golden():
    """Host half of the device-side collate (f2): the table padded ONCE with pad_token_lists and indexed by row equals
    what the reference pads per batch (its negative F.pad truncates long titles; empty titles become all-zero rows)."""
    from helpers import load_collate_golden
    from newsreclib_b200.data.components.device_collate import pad_token_lists
    news, splits, (lt, la) = load_collate_golden()
    title, abstract = pad_token_lists(news["tokenized_title"], lt), pad_token_lists(news["tokenized_abstract"], la)
    assert title.dtype == np.int64 and title.shape == (len(news["nid"]), lt)
    for split, (samples, ref) in splits.items():
        hist = np.concatenate([s[2] for s in samples])
        cand = np.concatenate([s[3] for s in samples])
        assert np.array_equal(title[hist], ref["x_hist"]["title"].numpy())
        assert np.array_equal(title[cand], ref["x_cand"]["title"].numpy())
        assert np.array_equal(abstract[hist], ref["x_hist"]["abstract"].numpy())
        assert np.array_equal(np.asarray(news["nid"])[cand], ref["x_cand"]["news_ids"].numpy())


@pytest.mark.parametrize("name", ["nrms_module_ref", "nrms_module_ref_late_fusion", "nrms_module_ref_supcon",
                                  "nrms_module_ref_dual"])
def test_oracle_matches_reference_nrms_module_golden(name):
    """The fixtures hold what the reference's OWN NRMSModule.forward (nrms_module.py:230-255) and model_step (:260-362)
    returned (oracle/make_module_golden.py; only to_dense_batch is the restated PyG function): the oracle's glue must
    reproduce scores, loss and gradients, and the 11-tuple must be what the GPU module's vectorised model_step assumes."""
    from helpers import grad_sample, load_module_golden
    params, batch, ref, meta = load_module_golden(name)
    lf = meta["late_fusion"]
    use = {k: v for k, v in params.items() if not (lf and k.startswith(USER))}
    lk = dict(loss_name=ref.get("loss_name", "cross_entropy_loss"), dual_loss_coef=ref.get("dual_loss_coef"))
    scores, loss, grads = oracle_run(use, batch, meta["H"], late_fusion=lf, **lk)
    assert scores.shape == ref["scores"].shape and rel_err(scores, ref["scores"]) <= 2e-6
    assert rel_err(loss, ref["out"]["loss"]) <= 2e-6
    tols = grad_tolerances(use, batch, meta["H"], 2e-5, grads, late_fusion=lf, **lk)
    for k, g in ref["grad"].items():
        if float(g.abs().max()) < 1e-9:
            continue  # mathematically zero gradient (key bias): rounding noise on both sides
        assert rel_err(grad_sample(grads[k]), g) <= tols[k], k
    # model_step's outputs in terms of the ragged batch (what two_tower.TwoTowerRecommender.model_step returns directly)
    out = ref["out"]
    B = meta["B"]
    sizes_c, sizes_h = torch.bincount(batch["batch_cand"], minlength=B), torch.bincount(batch["batch_hist"], minlength=B)
    assert torch.equal(out["cand_news_size"], sizes_c) and torch.equal(out["hist_news_size"], sizes_h)
    mask = torch.arange(ref["scores"].shape[1])[None, :] < sizes_c[:, None]
    assert torch.equal(out["preds"], ref["scores"][mask])                  # masked dense order == ragged order
    assert torch.equal(out["targets"], batch["labels"])
    assert torch.equal(out["target_categories"], batch["x_cand"]["category"])
    assert torch.equal(out["target_sentiments"], batch["x_cand"]["sentiment"])
    assert torch.equal(out["hist_categories"], batch["x_hist"]["category"])
    assert torch.equal(out["hist_sentiments"], batch["x_hist"]["sentiment"])
    assert torch.equal(out["user_ids"], batch["user_ids"]) and torch.equal(out["cand_news_ids"], batch["x_cand"]["news_ids"])
    assert bool((ref["scores"][~mask] == 0).all())                          # padded slots score exactly 0.0


@pytest.mark.parametrize("name", ["nrms_module_ref_supcon", "nrms_module_ref_dual"])
def test_sup_con_oracle_matches_reference_criterion_golden(name):
    """The SupCon / dual-loss fixtures hold the value and d loss / d scores of the reference's OWN criterion objects
    (components/losses.py over oracle/pml_standins.py, index tuples of nrms_module.py:290-307) on the reference's own
    scores: oracle.sup_con_loss must reproduce both from (scores, labels, segment ids) alone."""
    from helpers import load_module_golden
    _, batch, ref, meta = load_module_golden(name)
    s = ref["scores"].clone().requires_grad_(True)
    coef = ref["dual_loss_coef"] if ref["loss_name"] == "dual_loss" else None
    loss = O.nrms_loss(batch, s, ref["loss_name"], coef)
    loss.backward()
    assert rel_err(loss.detach(), ref["out"]["loss"]) <= 1e-6
    assert rel_err(s.grad, ref["d_scores"]) <= 1e-5
    sizes = torch.bincount(batch["batch_cand"], minlength=meta["B"])
    mask = torch.arange(s.shape[1])[None, :] < sizes[:, None]
    assert len(set(sizes.tolist())) > 1                                     # ragged: padded slots exist
    if ref["loss_name"] == "sup_con_loss":
        assert bool((ref["d_scores"][~mask] == 0).all())                    # padded slots are in neither index set
        y, _ = O.to_dense_batch(batch["labels"], batch["batch_cand"])
        no_pos = y.sum(dim=1) == 0
        assert bool(no_pos.any()) and bool((ref["d_scores"][no_pos] == 0).all())  # rows without a positive drop out


def test_sup_con_oracle_zero_loss_cases():
    """components/losses.py:14-15,20,40 -> zero_losses(): every index list has at most one element, or there is no
    positive / no negative in the whole batch; and AvgNonZeroReducer with no row > 0."""
    s = torch.randn(3, 4)
    m = torch.tensor([[1, 1, 0, 0], [1, 0, 0, 0], [1, 1, 1, 0]], dtype=torch.bool)
    assert float(O.sup_con_loss(s, torch.zeros(3, 4), m)) == 0.0                              # no positive at all
    y = torch.tensor([[1., 1, 0, 0], [1, 0, 0, 0], [1, 1, 1, 0]])
    assert float(O.sup_con_loss(s, y, m)) == 0.0                                              # no negative at all
    m1 = torch.tensor([[1, 1, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]], dtype=torch.bool)
    y1 = torch.tensor([[1., 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]])
    assert float(O.sup_con_loss(s, y1, m1)) == 0.0                                            # one positive, one negative pair
    y2 = torch.tensor([[1., 0, 0, 0], [0, 0, 0, 0], [0., 1, 0, 0]])
    s2 = s.clone().requires_grad_(True)
    l = O.sup_con_loss(s2, y2, m)
    assert float(l) > 0
    l.backward()
    assert bool((s2.grad[1] == 0).all()) and bool((s2.grad[~m] == 0).all())


def test_naml_fixture_is_what_the_reference_naml_module_returns():
    """naml_module_ref.npz = outputs of the reference's OWN NAMLModule.forward / model_step (naml_module.py:261-286 and
    model_step; oracle/make_module_golden.py) on the inputs of naml_mind.npz: the fixture the NAML parity tests use
    (minted from the reference's component modules + restated glue) must agree with it."""
    r = dict(np.load(os.path.join(GOLD, "naml_module_ref.npz")))
    g = dict(np.load(os.path.join(GOLD, str(r["source"]) + ".npz")))
    assert r["scores"].shape == g["scores"].shape and rel_err(r["scores"], g["scores"]) < 1e-6
    assert rel_err(r["loss"], g["loss"]) < 1e-6
    sizes = np.bincount(g["batch_cand"])
    mask = np.arange(r["scores"].shape[1])[None, :] < sizes[:, None]
    assert np.array_equal(r["preds"], r["scores"][mask]) and np.array_equal(r["targets"], g["labels"])
    assert np.array_equal(r["cand_news_size"], sizes) and np.array_equal(r["hist_news_size"], np.bincount(g["batch_hist"]))


def test_to_dense_batch_property_against_naive_loop():
    """PyG is absent (parity unpinned for this one third-party function), so the restatement is checked against the
    definition itself: item i of segment b lands at [b, rank of i within b], everything else is zero / False; segments
    may be empty, x may be 1-D (labels, nrms_module.py:277) or carry trailing dims."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(0, 6), min_size=1, max_size=9), st.sampled_from([(), (3,), (2, 2)]))
    def check(counts, trailing):
        if sum(counts) == 0 or counts[-1] == 0:
            counts = counts + [1]                      # B = batch.max() + 1: the last segment is never empty
        batch = torch.repeat_interleave(torch.arange(len(counts)), torch.tensor(counts))
        x = torch.arange(1, batch.numel() * int(np.prod(trailing, dtype=np.int64)) + 1, dtype=torch.float32)
        x = x.reshape((batch.numel(),) + trailing)
        dense, mask = O.to_dense_batch(x, batch)
        B, M = len(counts), max(counts)
        assert dense.shape == (B, M) + trailing and mask.shape == (B, M)
        want, wmask, i = torch.zeros((B, M) + trailing), torch.zeros(B, M, dtype=torch.bool), 0
        for b, c in enumerate(counts):
            for j in range(c):
                want[b, j], wmask[b, j] = x[i], True
                i += 1
        assert torch.equal(dense, want) and torch.equal(mask, wmask)
        assert torch.equal(dense[mask], x)             # masked dense order == ragged order (what model_step relies on)

    check()


# ---- transformer restatement against the fixtures minted from HF RobertaModel -------------
@pytest.mark.parametrize("name", ["tfm_tiny", "tfm_t40", "tfm_bert"])
def test_tfm_oracle_matches_hf_golden(name):
    from tfm_helpers import grad_errors, load_tfm_golden, oracle_tfm
    g, cfg, params, rgrads = load_tfm_golden(name)
    ids, att, w = torch.from_numpy(g["input_ids"]), torch.from_numpy(g["attention_mask"]), torch.from_numpy(g["w"])
    out, grads = oracle_tfm(params, cfg, ids, att, w, frozen=cfg["frozen"])
    assert rel_err(out, g["out"]) <= 5e-5
    assert set(grads) == set(rgrads)  # the frozen layers get none, the embeddings do
    errs = grad_errors(grads, rgrads)
    assert max(errs.values()) <= 5e-4, max(errs.items(), key=lambda kv: kv[1])
    # the padding rows never receive a gradient: RoBERTa nn.Embedding(padding_idx=1) for words AND positions; BERT
    # padding_idx=0 for words only (its position table has no padding row)
    pad = 0 if cfg["bert"] else 1
    assert float(grads["embeddings.word_embeddings.weight"][pad].abs().max()) == 0.0
    if not cfg["bert"]:
        assert float(grads["embeddings.position_embeddings.weight"][1].abs().max()) == 0.0
    else:
        assert float(grads["embeddings.position_embeddings.weight"][0].abs().max()) > 0.0


def test_tfm_oracle_position_ids_and_masking():
    from oracle import tfm_oracle as TO
    ids = torch.tensor([[0, 5, 6, 2, 1, 1], [0, 7, 2, 1, 1, 1]])
    assert TO.position_ids(ids, 1).tolist() == [[2, 3, 4, 5, 1, 1], [2, 3, 4, 1, 1, 1]]
    # a masked key has no influence on the valid positions
    from tfm_helpers import random_tfm_params
    P = random_tfm_params(64, 1, 128, 1, 20, 10, seed=0)
    att = (ids != 1).long()
    a = TO.encoder(ids, att, P, 1, 1)
    ids2 = ids.clone(); ids2[0, 4] = 9  # change a masked position's token (position id changes too)
    b = TO.encoder(ids2, att, P, 1, 1)
    assert torch.allclose(a[0, :4], b[0, :4], atol=1e-6) and not torch.allclose(a[0, 4], b[0, 4], atol=1e-3)


def test_plm_mirror_refuses_uncovered_transformers():
    """transformer_impl='native' (the default) never falls back silently: an architecture outside the kernels' coverage
    raises at construction and names the explicit opt-out."""
    from transformers import RobertaConfig, RobertaModel
    from newsreclib_b200.models.components.encoders.news.text import PLM
    kw = dict(frozen_layers=[0], use_mhsa=True, apply_reduce_dim=False, reduced_embed_dim=None, num_heads=2, query_dim=8,
              dropout_probability=0.2)
    small = RobertaModel(RobertaConfig(vocab_size=30, hidden_size=96, num_hidden_layers=1, num_attention_heads=2,
                                       intermediate_size=64, max_position_embeddings=20))
    with pytest.raises(ValueError, match="transformer_impl='hf'"):
        PLM(plm_model=small, embed_dim=96, **kw)
    assert PLM(plm_model=small, embed_dim=96, transformer_impl="hf", **kw).transformer_impl == "hf"
    ok = RobertaModel(RobertaConfig(vocab_size=30, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                    intermediate_size=64, max_position_embeddings=20))
    m = PLM(plm_model=ok, embed_dim=128, **kw)
    ps = m.transformer_parameters()
    assert len(ps) == 5 + 16 * 2 and ps[2].shape == (128,)
    assert [p.requires_grad for p in ps[5:21]] == [False] * 16 and all(p.requires_grad for p in ps[21:])
    with pytest.raises(RuntimeError, match="CUDA"):  # no CPU path
        m({"input_ids": torch.zeros(2, 5, dtype=torch.long), "attention_mask": torch.ones(2, 5, dtype=torch.long)})


def test_transformer_hidden_states_do_not_depend_on_trailing_padding():
    """The claim PLM.forward_pair rests on (one transformer pass for history + candidates, each part cut back to its own
    padded length): with the key-padding mask, a token's hidden state is the same whether its text is padded to 9 or to 14
    positions -- checked on the real HF RobertaModel and on the restatement."""
    from transformers import RobertaConfig, RobertaModel
    from oracle import tfm_oracle as TO
    torch.manual_seed(3)
    tf = RobertaModel(RobertaConfig(vocab_size=50, hidden_size=64, num_hidden_layers=2, num_attention_heads=1,
                                    intermediate_size=96, max_position_embeddings=20, pad_token_id=1, type_vocab_size=1),
                      add_pooling_layer=False).eval()
    ids = torch.tensor([[0, 7, 9, 12, 2, 1, 1, 1, 1], [0, 5, 6, 7, 8, 9, 10, 11, 2]])
    att = (ids != 1).long()
    wide_ids = torch.nn.functional.pad(ids, (0, 5), value=1)
    wide_att = torch.nn.functional.pad(att, (0, 5), value=0)
    with torch.no_grad():
        a = tf(input_ids=ids, attention_mask=att)[0]
        b = tf(input_ids=wide_ids, attention_mask=wide_att)[0][:, :9]
    assert float((a - b).abs().max()) <= 2e-6  # padding positions included: they attend to the same valid keys
    P = {k: v for k, v in tf.state_dict().items() if v.is_floating_point()}
    eps = tf.config.layer_norm_eps
    oa = TO.encoder(ids, att, P, 1, 2, eps=eps)
    ob = TO.encoder(wide_ids, wide_att, P, 1, 2, eps=eps)[:, :9]
    assert float((oa - ob).abs().max()) <= 2e-6 and float((oa - a).abs().max()) <= 2e-5


def test_two_tower_merges_per_news_encoder_calls():
    """TwoTowerRecommender._encode_hist_and_cand: ONE call of a per-news encoder over history + candidates gives the vectors
    of the two separate calls (nrms_module.py:231,235); mismatching token widths fall back to two calls."""
    from newsreclib_b200.models.general_rec.two_tower import TwoTowerRecommender

    class ToyNews(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.text_encoders = torch.nn.ModuleDict({"title": torch.nn.Embedding(30, 8)})
            self.calls = 0

        def forward(self, news):
            self.calls += 1
            return self.text_encoders["title"](news["title"]).mean(dim=1)  # per-news: rows independent
    m = TwoTowerRecommender.__new__(TwoTowerRecommender)
    torch.nn.Module.__init__(m)
    m.news_encoder = ToyNews()
    h, c = {"title": torch.randint(0, 30, (7, 5))}, {"title": torch.randint(0, 30, (3, 5))}
    vh, vc = m._encode_hist_and_cand(h, c)
    assert m.news_encoder.calls == 1 and vh.shape == (7, 8) and vc.shape == (3, 8)
    assert torch.equal(vh, m.news_encoder(h)) and torch.equal(vc, m.news_encoder(c))
    m.news_encoder.calls = 0
    m._encode_hist_and_cand(h, {"title": torch.randint(0, 30, (3, 6))})  # another padded width: two calls
    assert m.news_encoder.calls == 2
    m.merge_news_calls, m.news_encoder.calls = False, 0
    m._encode_hist_and_cand(h, c)
    assert m.news_encoder.calls == 2
