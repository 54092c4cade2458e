"""GPU tests of the rows either side of the encoders: device-side collate against the
restated reference collate (bit-exact, integer / index work), and the cached-news-vector evaluation path
against the plain forward."""
import numpy as np
import pytest
import torch

from helpers import TITLE, rel_err
from newsreclib_b200.synthetic import make_nrms_params, make_titles

pytestmark = pytest.mark.gpu


def make_news(rng, M, V, with_abstract=True):
    lens = rng.integers(0, 45, M)
    news = {
        "nid": [int(x) for x in rng.integers(1, 10 ** 6, M)],
        "tokenized_title": [[int(t) for t in rng.integers(1, V, int(l))] for l in lens],  # some empty, some > 30
        "category_class": [int(x) for x in rng.integers(1, 19, M)],
        "subcategory_class": [int(x) for x in rng.integers(1, 200, M)],
        "sentiment_class": [int(x) for x in rng.integers(1, 4, M)],
        "sentiment_score": [float(x) for x in rng.random(M)],
    }
    if with_abstract:
        news["tokenized_abstract"] = [[int(t) for t in rng.integers(1, V, int(l))] for l in rng.integers(0, 80, M)]
    return news


def make_samples(rng, B, M):
    out = []
    for b in range(B):
        h = rng.integers(0, M, int(rng.integers(1, 51)))
        c = rng.integers(0, M, int(rng.integers(2, 40)))
        y = np.zeros(len(c), dtype=np.int64); y[rng.integers(0, len(c))] = 1
        out.append((np.array([int(rng.integers(1, 10 ** 6))]), np.array([b]), h, c, y))
    return out


def test_device_collate_matches_reference_collate_bit_exact():
    from newsreclib_b200.data.components.device_collate import DeviceCollate, DeviceNewsTable
    from oracle import collate_oracle as CO
    rng = np.random.default_rng(3)
    M, V = 500, 3000
    news = make_news(rng, M, V)
    table = DeviceNewsTable.from_token_lists(
        news["nid"], news["tokenized_title"], news["category_class"], news["subcategory_class"], 30,
        news["tokenized_abstract"], 50, news["sentiment_class"], news["sentiment_score"])
    samples = make_samples(rng, 16, M)
    got = DeviceCollate(table)(samples)
    ref = CO.collate(news, samples, 30, 50)
    for k in ("batch_hist", "batch_cand", "labels", "user_ids", "user_idx"):
        assert got[k].dtype == ref[k].dtype and torch.equal(got[k].cpu(), ref[k]), k
    for side in ("x_hist", "x_cand"):
        assert set(got[side]) == set(ref[side])
        for k, v in ref[side].items():
            assert got[side][k].dtype == v.dtype and torch.equal(got[side][k].cpu(), v), (side, k)
    assert got["x_hist"]["title"].shape[1] == 30 and got["x_hist"]["abstract"].shape[1] == 50
    with pytest.raises(KeyError):
        DeviceCollate(table)([(np.array([1]), np.array([0]), np.array([M]), np.array([0, 1]), np.array([1, 0]))])


def test_gather_rows_word_sizes():
    from newsreclib_b200 import ops
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, 97, (1000,), generator=g).cuda()
    for t in (torch.randint(0, 10 ** 9, (97, 30), generator=g), torch.randint(0, 99, (97,), generator=g),
              torch.randn(97, generator=g), torch.randn(97, 301, generator=g), torch.randn(97, 300, generator=g)):
        tc = t.cuda()
        assert torch.equal(ops.gather_rows(tc, idx), tc[idx])


def test_cached_news_vectors_eval_path_matches_forward():
    from test_gpu_modules import full_batch, make_module
    from newsreclib_b200.synthetic import make_batch
    V, M, B = 2000, 700, 12
    params = make_nrms_params(V, seed=4)
    rng = np.random.default_rng(4)
    titles = torch.from_numpy(make_titles(rng, M, V, 30))
    m = make_module(params).cuda().eval()
    m.load_state_dict(params)
    vecs = m.encode_news_table({"title": titles.cuda()}, chunk=256)
    assert vecs.shape == (M, 300)
    # impressions referencing table rows; the plain forward sees the gathered titles
    batch = make_batch(B, V, hist="ragged", cand="eval", seed=9, max_hist=20)
    nh, nc = batch["batch_hist"].numel(), batch["batch_cand"].numel()
    hist_rows = torch.from_numpy(rng.integers(0, M, nh)); cand_rows = torch.from_numpy(rng.integers(0, M, nc))
    batch["x_hist"]["title"] = titles[hist_rows]; batch["x_cand"]["title"] = titles[cand_rows]
    b = full_batch(batch)
    ref = m(b)
    got = m.forward_cached(vecs, hist_rows.cuda(), b["batch_hist"], cand_rows.cuda(), b["batch_cand"], B)
    assert got.shape == ref.shape
    assert rel_err(got, ref) <= 1e-6  # same kernels on the same rows: identical up to launch geometry
    # ranking metrics on device from the cached scores
    from newsreclib_b200.metrics import ranking_metrics
    sizes = torch.bincount(batch["batch_cand"])
    mask = torch.arange(got.shape[1], device="cuda")[None, :] < sizes.cuda()[:, None]
    met = ranking_metrics(got[mask], b["labels"], sizes.cuda(), [5, 10])
    assert 0.0 <= float(met["auc"]) <= 1.0 and 0.0 < float(met["mrr"]) <= 1.0


def test_cached_eval_path_full_impression_size_vs_oracle():
    """SURVEY section 8 f4 at evaluation size: 64 whole impressions (candidate lists up to 300, histories up to 50) scored
    from news vectors encoded ONCE for the news table, against the CPU oracle's eval forward on the gathered titles
    (logit bar 1e-4), padded slots exactly 0.0, and the device ranking metrics against their definitions on the CPU."""
    from test_gpu_modules import full_batch, make_module
    from test_gpu_fullsize import _eval_batch_with_long_impression
    from newsreclib_b200.metrics import ranking_metrics
    from oracle import nrms_oracle as O
    V, M, B = 70000, 6000, 64
    params = make_nrms_params(V, seed=44)
    rng = np.random.default_rng(44)
    titles = torch.from_numpy(make_titles(rng, M, V, 30))
    m = make_module(params).cuda().eval()
    m.load_state_dict(params)
    vecs = m.encode_news_table({"title": titles.cuda()})
    batch = _eval_batch_with_long_impression(B, V, 300, seed=45)
    nh, nc = batch["batch_hist"].numel(), batch["batch_cand"].numel()
    hist_rows = torch.from_numpy(rng.integers(0, M, nh)); cand_rows = torch.from_numpy(rng.integers(0, M, nc))
    batch["x_hist"]["title"] = titles[hist_rows]; batch["x_cand"]["title"] = titles[cand_rows]
    got = m.forward_cached(vecs, hist_rows.cuda(), batch["batch_hist"].cuda(), cand_rows.cuda(), batch["batch_cand"].cuda(), B)
    with torch.no_grad():
        ref = O.nrms_forward(batch, params, 15)
    assert got.shape == ref.shape == (B, 300)
    e = rel_err(got, ref)
    print(f"cached eval path, B=64 Cmax=300: logits rel {e:.2e} (tol 1e-4)")
    assert e <= 1e-4
    sizes = torch.bincount(batch["batch_cand"], minlength=B)
    for b in range(B):
        assert torch.all(got[b, sizes[b]:] == 0)
    mask = torch.arange(300)[None, :] < sizes[:, None]
    met = ranking_metrics(got[mask.cuda()], batch["labels"].cuda(), sizes.cuda(), [5, 10])
    # the same definitions, impression by impression, on the oracle's scores
    mrr, nd5 = [], []
    off = np.concatenate([[0], np.cumsum(sizes.numpy())])
    for b in range(B):
        sc, y = ref[b, :sizes[b]].numpy(), batch["labels"][off[b]:off[b + 1]].numpy()
        order = np.argsort(-sc, kind="stable")
        ys = y[order]
        mrr.append(1.0 / (np.argmax(ys > 0) + 1) if ys.sum() > 0 else 0.0)
        disc = 1.0 / np.log2(np.arange(2, len(ys) + 2))
        idcg = (np.sort(y)[::-1][:5] * disc[:5]).sum()
        nd5.append((ys[:5] * disc[:5]).sum() / idcg if idcg > 0 else 0.0)
    assert abs(float(met["mrr"]) - np.mean(mrr)) < 2e-3 and abs(float(met["ndcg@5"]) - np.mean(nd5)) < 2e-3


def test_test_epoch_hooks_aspect_metrics_and_recommendation_dump(tmp_path):
    """The reference's test stage (nrms_module.py:442-535): test_step over two batches, then on_test_epoch_end returns the
    ranking metrics, categ/sent diversity and personalization @k (against a per-impression python loop over the module's
    own scores) and, with save_recs=True, writes {"U<user>": {"N<news>": score}} (abstract_recommender.py:150-185)."""
    import functools
    import json
    from helpers import TITLE
    from test_gpu_modules import OUTPUTS, full_batch
    from newsreclib_b200.models.general_rec.nrms_module import NRMSModule
    from newsreclib_b200.synthetic import make_batch
    V = 1500
    params = make_nrms_params(V, seed=5)
    outputs = dict(OUTPUTS)
    outputs["test"] = OUTPUTS["test"] + ["target_categories", "target_sentiments", "hist_categories", "hist_sentiments"]
    path = str(tmp_path / "recs.json")
    m = NRMSModule(
        dataset_attributes=["title", "category", "sentiment"], attributes2encode=["title"], outputs=outputs,
        dual_loss_training=False, dual_loss_coef=None, loss="cross_entropy_loss", late_fusion=False, temperature=None,
        use_plm=False, pretrained_embeddings_path=None, plm_model=None, frozen_layers=None, embed_dim=300, num_heads=15,
        query_dim=200, dropout_probability=0.2, top_k_list=[5, 10], num_categ_classes=18, num_sent_classes=3,
        save_recs=True, recs_fpath=path, optimizer=functools.partial(torch.optim.Adam, lr=1e-4), scheduler=None,
        pretrained_embeddings=params[TITLE + "embedding_layer.weight"])
    m.load_state_dict(params)
    m = m.cuda().eval()
    batches = [make_batch(7, V, hist="ragged", cand="eval", seed=s, max_hist=12) for s in (71, 72)]
    with torch.no_grad():
        for i, b in enumerate(batches):
            m.test_step(full_batch(b), i)
        scores = [m(full_batch(b)).cpu() for b in batches]
    met = m.on_test_epoch_end()
    for k in ("test/auc", "test/mrr", "test/ndcg@5", "test/categ_div@5", "test/categ_pers@10", "test/sent_div@10",
              "test/sent_pers@5"):
        assert k in met and 0.0 <= float(met[k]) <= 1.0, k
    # per-impression loop over the same scores (definitions of metrics/functional.py:8-49,52-110)
    div5, pers10, recs = [], [], {}
    for b, sc in zip(batches, scores):
        B = 7
        cs, hs = torch.bincount(b["batch_cand"], minlength=B), torch.bincount(b["batch_hist"], minlength=B)
        co, ho = np.concatenate([[0], np.cumsum(cs.numpy())]), np.concatenate([[0], np.cumsum(hs.numpy())])
        for i in range(B):
            s = sc[i, :cs[i]].numpy()
            cat = b["x_cand"]["category"][co[i]:co[i + 1]].numpy()
            order = np.argsort(-s, kind="stable")
            c5 = np.bincount(cat[order][:5], minlength=19).astype(np.float64)
            p = c5 / c5.sum()
            div5.append(float(-(p[p > 0] * np.log(p[p > 0])).sum() / np.log(19)))
            c10 = np.bincount(cat[order][:10], minlength=19)
            h = np.bincount(b["x_hist"]["category"][ho[i]:ho[i + 1]].numpy(), minlength=19)
            pers10.append(float(np.minimum(c10, h).sum() / np.maximum(c10, h).sum()))
            d = recs.setdefault(f"U{int(b['user_ids'][i])}", {})
            for n, v in zip(b["x_cand"]["news_ids"][co[i]:co[i + 1]].tolist(), s.tolist()):
                d[f"N{n}"] = v
    assert abs(float(met["test/categ_div@5"]) - np.mean(div5)) < 1e-5
    assert abs(float(met["test/categ_pers@10"]) - np.mean(pers10)) < 1e-5
    got = json.load(open(path))
    assert set(got) == set(recs)
    for u in recs:
        assert set(got[u]) == set(recs[u])
        assert max(abs(got[u][n] - recs[u][n]) for n in recs[u]) < 1e-5
    assert all(len(v) == 0 for v in m.test_step_outputs.values())          # cleared for the next epoch


@pytest.mark.parametrize("split", ["1", "0"])
def test_train_step_from_host_buffers_matches_device_resident_step(split, monkeypatch):
    """The end-to-end call (host buffers in, scores + loss out): nrl_nrms_step_host_begin / _end with the optimizer step
    queued between them (NRL_E2E_SPLIT=1, the default) and the one-call nrl_nrms_step_host (=0) against the
    device-resident train_step on the same seeds: scores, loss and the parameters after three steps; then a token id
    outside the table must come back as an error from the SAME call (device-side check word copied with the results)."""
    from helpers import batch_sizes, to_dev
    from newsreclib_b200 import ops
    from newsreclib_b200.synthetic import make_batch
    from newsreclib_b200.trainer import NRMSTrainer
    monkeypatch.setenv("NRL_E2E_SPLIT", split)
    V = 3000
    params = make_nrms_params(V, seed=8)
    batches = [make_batch(9, V, hist="ragged", cand="train", seed=80 + i, max_hist=14) for i in range(3)]
    t_dev = NRMSTrainer(params, 15, dropout_p=0.2, seed=5)
    t_host = NRMSTrainer(params, 15, dropout_p=0.2, seed=5)
    pin = lambda b: {k: ({c: t.pin_memory() for c, t in v.items()} if isinstance(v, dict) else v.pin_memory())
                     for k, v in to_dev(b, "cpu").items()}
    for b in batches:
        B, Hmax, Cmax = batch_sizes(b)
        s_ref, l_ref = t_dev.train_step(to_dev(b), B, Hmax, Cmax)
        scores_host, loss_host = torch.empty(B, Cmax).pin_memory(), torch.empty(1).pin_memory()
        t_host.train_step_host(pin(b), B, Hmax, Cmax, scores_host, loss_host)
        # results are on the host when the call returns: no synchronize here on purpose
        assert rel_err(scores_host, s_ref.cpu()) <= 1e-5 and rel_err(loss_host, l_ref.cpu()) <= 1e-5
    torch.cuda.synchronize()
    # the two replicas differ by the summation order of the gradient atomics; Adam turns a gradient that is pure rounding
    # noise (mathematically zero: the key third of in_proj_bias) into a +-lr move, so the criterion is the one of
    # NRMSTrainer.probe_exchange: nearly every element identical to fp32 rounding, none further apart than 3 steps of lr
    d = (t_host.flat - t_dev.flat).abs()
    assert float((d > 1e-6).float().mean()) < 2e-3 and float(d.median()) < 1e-7 and float(d.max()) <= 3.5e-4
    assert t_host.step_count == 3 and ops.device_status(raise_on_error=False) == 0
    bad = pin(batches[0])
    bad["x_cand"]["title"][0, 0] = V + 11
    B, Hmax, Cmax = batch_sizes(batches[0])
    with pytest.raises(RuntimeError, match="token id"):
        t_host.train_step_host(bad, B, Hmax, Cmax, torch.empty(B, Cmax).pin_memory(), torch.empty(1).pin_memory())
    assert ops.device_status(raise_on_error=False) == 0                      # reported once, then clear


@pytest.mark.parametrize("quantise,long_imp", [(False, False), (True, False), (False, True)])
def test_rank_metrics_kernel_matches_host_definitions(quantise, long_imp):
    """nrl_rank_metrics (one CTA per impression, ranks counted in shared memory) against the torch formulation of the same
    definitions on the host (metrics.ranking_metrics on CPU tensors: the one tests/test_metrics_cpu.py pins against
    scikit-learn): tied scores (stable order), impressions without a positive, a single candidate, an impression longer
    than the shared-memory tile, and the per-candidate ranks against a stable argsort."""
    from newsreclib_b200 import ops
    from newsreclib_b200.metrics import ranking_metrics
    g = torch.Generator().manual_seed(11 + int(quantise))
    sizes = torch.randint(1, 300, (50,), generator=g)
    sizes[3] = 1
    if long_imp:
        sizes[7] = 2500
    N = int(sizes.sum())
    preds = torch.randn(N, generator=g)
    if quantise:
        preds = (preds * 2).round() / 2
    targets = (torch.rand(N, generator=g) < 0.15).float()
    off = torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)])
    targets[off[5]:off[6]] = 0                                            # an impression without a positive
    ks = [1, 5, 10, 400]
    want = ranking_metrics(preds, targets, sizes, ks)
    got = ranking_metrics(preds.cuda(), targets.cuda(), sizes.cuda(), ks)
    for k in want:
        assert abs(float(got[k]) - float(want[k])) <= 2e-6, (k, float(got[k]), float(want[k]))
    per, ranks = ops.rank_metrics(preds.cuda(), targets.cuda(), sizes.cuda(), ks, want_ranks=True)
    assert per.shape == (50, 5) and float(per[5].abs().max()) == 0.0
    ranks = ranks.cpu()
    for b in (0, 3, 7, 20):
        s = preds[off[b]:off[b + 1]]
        order = torch.argsort(s, descending=True, stable=True)
        expect = torch.empty_like(order)
        expect[order] = torch.arange(1, s.numel() + 1)
        assert torch.equal(ranks[off[b]:off[b + 1]].long(), expect)
