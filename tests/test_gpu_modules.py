"""GPU tests of the drop-in boundary: the reference-shaped modules (same constructor kwargs and
state_dict keys) driven like the reference drives them, against the CPU oracle."""
import functools

import pytest
import torch

from helpers import TITLE, grad_tolerances, oracle_run, rel_err, to_dev
from newsreclib_b200.synthetic import make_batch, make_nrms_params

pytestmark = pytest.mark.gpu

OUTPUTS = {"train": ["preds", "targets", "cand_news_size"], "val": ["preds", "targets", "cand_news_size"],
           "test": ["preds", "targets", "cand_news_size", "hist_news_size", "user_ids", "cand_news_ids"]}


def make_module(params, late_fusion=False, p=0.2, loss="cross_entropy_loss", dual_loss_coef=None):
    from newsreclib_b200.models.general_rec.nrms_module import NRMSModule
    m = NRMSModule(
        dataset_attributes=["title", "category"], attributes2encode=["title"], outputs=OUTPUTS,
        dual_loss_training=loss == "dual_loss", dual_loss_coef=dual_loss_coef, loss=loss, late_fusion=late_fusion,
        temperature=None, use_plm=False, pretrained_embeddings_path=None, plm_model=None, frozen_layers=None,
        embed_dim=300, num_heads=15, query_dim=200, dropout_probability=p, top_k_list=[5, 10],
        num_categ_classes=18, num_sent_classes=3, save_recs=False, recs_fpath=None,
        optimizer=functools.partial(torch.optim.Adam, lr=1e-4), scheduler=None,
        pretrained_embeddings=params[TITLE + "embedding_layer.weight"])
    return m


def full_batch(batch, dev="cuda"):
    b = to_dev(batch, dev)
    for side in ("x_hist", "x_cand"):
        for k in ("category", "sentiment", "news_ids"):
            b[side][k] = batch[side][k].to(dev)
    b["user_ids"] = batch["user_ids"].to(dev)
    b["user_idx"] = batch["user_idx"].to(dev)
    return b


def test_nrms_module_forward_backward_matches_oracle():
    V = 2500
    params = make_nrms_params(V, seed=21)
    batch = make_batch(12, V, hist="ragged", cand="train", seed=21, max_hist=15)
    m = make_module(params)
    missing = m.load_state_dict(params, strict=True)  # reference key names load as-is
    assert not missing.missing_keys and not missing.unexpected_keys
    m = m.cuda().eval()
    b = full_batch(batch)
    scores = m(b)
    rs, rl, rg = oracle_run(params, batch, 15)
    assert rel_err(scores, rs) <= 1e-4
    out = m.model_step(b)
    assert len(out) == 11
    loss, preds, targets, cand_size, hist_size = out[:5]
    assert rel_err(loss, rl) <= 1e-4
    assert torch.equal(cand_size.cpu(), torch.bincount(batch["batch_cand"]))
    assert torch.equal(hist_size.cpu(), torch.bincount(batch["batch_hist"]))
    assert preds.numel() == batch["labels"].numel() and torch.equal(targets.cpu(), batch["labels"])
    loss.backward()
    tols = grad_tolerances(params, batch, 15, 1e-3, rg)
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        assert rel_err(p.grad, rg[k]) <= tols[k], k
    assert float(m.news_encoder.text_encoders["title"].embedding_layer.weight.grad[0].abs().max()) == 0.0


def test_nrms_module_trains_and_roundtrips_state_dict():
    V = 1500
    params = make_nrms_params(V, seed=5)
    m = make_module(params).cuda().train()
    m.load_state_dict(params)
    opt = m.configure_optimizers()["optimizer"]
    losses = []
    batch = full_batch(make_batch(8, V, hist="ragged", seed=5, max_hist=10))
    for step in range(8):
        opt.zero_grad()
        loss = m.training_step(batch, step)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(l == l for l in losses) and losses[-1] < losses[0]  # finite and overfitting one batch
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    m2 = make_module(params).cuda().eval()
    m2.load_state_dict(sd)
    m.eval()
    assert torch.equal(m(batch), m2(batch))  # deterministic in eval mode
    metrics = m.on_train_epoch_end()
    assert {"train/auc", "train/mrr", "train/ndcg@5", "train/ndcg@10"} <= set(metrics)


def test_late_fusion_module():
    V = 1200
    params = make_nrms_params(V, seed=8)
    batch = make_batch(5, V, hist="ragged", seed=8, max_hist=6)
    m = make_module(params, late_fusion=True)
    m.load_state_dict({k: v for k, v in params.items() if k.startswith(TITLE)}, strict=True)
    m = m.cuda().eval()
    rs, _, _ = oracle_run(params, batch, 15, late_fusion=True, grad=False)
    assert rel_err(m(full_batch(batch)), rs) <= 1e-4


def test_component_interfaces():
    from newsreclib_b200.models.components.layers.attention import AdditiveAttention
    from newsreclib_b200.models.components.layers.click_predictor import DotProduct
    from oracle import nrms_oracle as O
    with pytest.raises(ValueError):
        AdditiveAttention(input_dim=300.0, query_dim=200)
    from newsreclib_b200.models.components.encoders.news.text import MHSAAddAtt
    with pytest.raises(ValueError):
        MHSAAddAtt(torch.randn(10, 300), 300, 15, 200, 1)  # dropout_probability must be a float
    torch.manual_seed(0)
    add = AdditiveAttention(400, 200).cuda()
    x = torch.randn(7, 3, 400).cuda()  # NAML view combiner shape (news.py:162-163)
    with torch.no_grad():
        ref = O.additive_attention(x.cpu(), add.linear.weight.cpu(), add.linear.bias.cpu(), add.query.cpu())
    assert rel_err(add(x), ref) <= 1e-4
    u, c = torch.randn(6, 1, 300).cuda(), torch.randn(6, 300, 9).cuda()
    assert rel_err(DotProduct()(u, c), torch.bmm(u, c).squeeze(1)) <= 1e-5


@pytest.mark.parametrize("late_fusion", [False, True])
def test_fused_model_step_node_matches_per_op_functions(late_fusion):
    """NRMSModule.model_step as ONE autograd node (ops.NrmsStepFn: nrl_nrms_step forward, nrl_nrms_step_bwd backward) against
    the same module on the per-op autograd.Functions: eval mode and train mode (same CPU-RNG seed -> same keep-bit words),
    a scaled objective (d objective / d loss is read on the device), direct accumulation into preallocated .grad
    buffers (what ModuleTrainer uses), and a second backward refusing loudly."""
    V = 1800
    params = make_nrms_params(V, seed=13)
    use = {k: v for k, v in params.items() if not (late_fusion and not k.startswith(TITLE))}
    batch = full_batch(make_batch(10, V, hist="ragged", cand="train", seed=13, max_hist=12))

    def run(fused, train, scale=1.0, direct=False):
        m = make_module(params, late_fusion=late_fusion)
        m.load_state_dict(use, strict=True)
        m = m.cuda()
        m.train(train)
        m.fused_model_step = fused
        if direct:
            for p in m.parameters():
                p.grad = torch.full_like(p, 2.0 ** -12)    # accumulation (+=) into what is already there
            m.grad_targets = "param.grad"
        torch.manual_seed(3)
        out = m.model_step(batch)
        (out[0] * scale).backward()
        grads = {k: p.grad.clone() - (2.0 ** -12 if direct else 0.0) for k, p in m.named_parameters()}
        return m, out, grads

    for train in (True, False):
        _, ref_out, ref_g = run(False, train)
        m, out, g = run(True, train)
        assert len(out) == 11 and not out[1].requires_grad
        assert rel_err(out[0].detach(), ref_out[0].detach()) <= 1e-5 and rel_err(out[1], ref_out[1].detach()) <= 1e-5
        worst = max(rel_err(g[k], ref_g[k]) for k in ref_g if float(ref_g[k].abs().max()) > 1e-9)
        print(f"fused model_step vs per-op Functions (late_fusion={late_fusion}, train={train}): worst gradient rel {worst:.2e} (tol 2e-4)")
        assert worst <= 2e-4
        with pytest.raises(RuntimeError, match="backward called twice"):
            out[0].backward()
    _, _, g_half = run(True, False, scale=0.5)
    _, _, g_direct = run(True, False, direct=True)
    for k in ref_g:
        if float(ref_g[k].abs().max()) > 1e-9:
            assert rel_err(2.0 * g_half[k], g[k]) <= 1e-5, k
            # every partial sum the kernels add into the prefilled 2^-12 is rounded at 2^-36: an absolute floor on top
            assert float((g_direct[k] - g[k]).abs().max()) <= 1e-5 * float(g[k].abs().max()) + 1e-9, k


def test_optim_adam_matches_torch_adam():
    """newsreclib_b200.optim.Adam (one nrl_adam_step launch per parameter) against torch.optim.Adam over five steps of the
    same gradients: parameters and both moments (configs/model/nrms.yaml:49-52; abstract_recommender.py:89-108), torch's
    state_dict key names, and the options it does not build refusing loudly."""
    from newsreclib_b200.optim import Adam
    g = torch.Generator().manual_seed(0)
    shapes = [(700, 300), (900, 300), (900,), (200,), (3,)]
    p_ref = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    p_new = [torch.nn.Parameter(p.detach().clone()) for p in p_ref]
    o_ref, o_new = torch.optim.Adam(p_ref, lr=1e-3), Adam(p_new, lr=1e-3)
    for step in range(5):
        for a, b in zip(p_ref, p_new):
            grad = torch.randn(a.shape, generator=g).cuda() * (0.0 if (step == 2 and a.dim() == 1) else 1.0)
            a.grad, b.grad = grad.clone(), grad.clone()
        o_ref.step(); o_new.step()
    for a, b in zip(p_ref, p_new):
        assert rel_err(b.detach(), a.detach()) <= 2e-6
        assert rel_err(o_new.state[b]["exp_avg"], o_ref.state[a]["exp_avg"]) <= 2e-6
        assert rel_err(o_new.state[b]["exp_avg_sq"], o_ref.state[a]["exp_avg_sq"]) <= 2e-6
        assert int(o_new.state[b]["step"]) == 5
    assert set(o_new.state_dict()["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    with pytest.raises(NotImplementedError):
        Adam(p_new, weight_decay=0.01)
