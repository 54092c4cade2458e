"""GPU tests against fixtures that hold the output of the reference's OWN top-level code, run in the build container
under stand-ins for its absent third-party imports: ``NRMSModule.forward`` / ``model_step``
(``oracle/make_module_golden.py``) and ``DatasetCollate`` (``oracle/make_collate_golden.py``)."""
import pytest
import torch

from helpers import grad_tolerances, oracle_run, rel_err
from test_gpu_modules import full_batch, make_module

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["nrms_module_ref", "nrms_module_ref_late_fusion", "nrms_module_ref_supcon",
                                  "nrms_module_ref_dual"])
def test_nrms_module_matches_reference_module_golden(name):
    """The drop-in NRMSModule against what the reference's OWN NRMSModule.forward / model_step returned on the same
    batch and weights (tests/golden/nrms_module_ref*.npz, minted by oracle/make_module_golden.py): scores, loss, every
    entry of the 11-tuple, and the gradients."""
    from helpers import MODULE_OUT, USER, grad_sample, load_module_golden
    params, batch, ref, meta = load_module_golden(name)
    lf = meta["late_fusion"]
    use = {k: v for k, v in params.items() if not (lf and k.startswith(USER))}
    lk = dict(loss_name=ref.get("loss_name", "cross_entropy_loss"), dual_loss_coef=ref.get("dual_loss_coef"))
    if lk["loss_name"] != "dual_loss":
        lk["dual_loss_coef"] = None
    m = make_module(params, late_fusion=lf, loss=lk["loss_name"], dual_loss_coef=lk["dual_loss_coef"])
    m.load_state_dict(use, strict=True)
    m = m.cuda().eval()
    b = full_batch(batch)
    scores = m(b)
    assert scores.shape == ref["scores"].shape and rel_err(scores, ref["scores"]) <= 1e-4
    B = meta["B"]
    sizes_c = torch.bincount(batch["batch_cand"], minlength=B)
    mask = torch.arange(scores.shape[1])[None, :] < sizes_c[:, None]
    assert bool((scores.cpu()[~mask] == 0).all())                           # padded slots exactly 0.0
    out = m.model_step(b)
    assert len(out) == len(MODULE_OUT)
    got = dict(zip(MODULE_OUT, out))
    assert rel_err(got["loss"], ref["out"]["loss"]) <= 1e-4
    assert got["preds"].shape == ref["out"]["preds"].shape and rel_err(got["preds"], ref["out"]["preds"]) <= 1e-4
    for k in MODULE_OUT[2:]:                                                # integer / label outputs: exact
        assert torch.equal(got[k].cpu().to(ref["out"][k].dtype), ref["out"][k]), k
    got["loss"].backward()
    _, _, rg = oracle_run(use, batch, meta["H"], late_fusion=lf, **lk)
    tols = grad_tolerances(use, batch, meta["H"], 1e-3, rg, late_fusion=lf, **lk)
    worst = max(((rel_err(grad_sample(p.grad), ref["grad"][k]) / (2.0 * tols[k]), k) for k, p in m.named_parameters()
                 if float(ref["grad"][k].abs().max()) >= 1e-9), default=(0.0, ""))
    print(f"{name}: loss rel {rel_err(got['loss'], ref['out']['loss']):.2e} (tol 1e-4); worst gradient at "
          f"{worst[0]:.2f} of its tolerance ({worst[1]}, tol {2.0 * tols.get(worst[1], 0):.1e})")
    for k, p in m.named_parameters():
        g = ref["grad"][k]
        if float(g.abs().max()) < 1e-9:
            continue  # mathematically zero gradient (key bias): rounding noise on both sides
        assert p.grad is not None, k
        # 2 x: the stored sample of a big gradient normalises by the sample's maximum, not the tensor's
        assert rel_err(grad_sample(p.grad), g) <= 2.0 * tols[k], k


@pytest.mark.parametrize("name", ["nrms_module_ref_supcon", "nrms_module_ref_dual"])
def test_sup_con_kernels_match_reference_criterion_golden(name):
    """nrl_supcon_fwd / nrl_supcon_bwd (+ nrl_ce_soft_* for the dual loss) on the reference's own scores against the value
    and d loss / d scores of the reference's OWN criterion objects (components/losses.py:6-40 with the index tuples of
    nrms_module.py:290-307; oracle/make_module_golden.py)."""
    from helpers import load_module_golden
    from newsreclib_b200 import ops
    _, batch, ref, meta = load_module_golden(name)
    s = ref["scores"].cuda().requires_grad_(True)
    off = ops.segment_offsets(batch["batch_cand"].cuda(), meta["B"])
    coef = ref["dual_loss_coef"] if ref["loss_name"] == "dual_loss" else None
    loss = ops.SupConFn.apply(s, batch["labels"].cuda(), off, coef)
    loss.backward()
    el, eg = rel_err(loss.detach(), ref["out"]["loss"]), rel_err(s.grad, ref["d_scores"])
    print(f"{name}: loss rel {el:.2e} (tol 1e-5)  d_scores rel {eg:.2e} (tol 1e-4)")
    assert el <= 1e-5 and eg <= 1e-4
    if coef is None:
        assert bool((s.grad.cpu()[ref["d_scores"] == 0] == 0).all())        # padded slots / rows without a positive: exact 0


@pytest.mark.parametrize("B,cmax,case", [(70, 300, "ragged"), (3, 9, "ragged"), (40, 5, "train"), (5, 6, "no_pos"),
                                         (4, 7, "no_neg"), (2, 1, "one_pair")])
def test_sup_con_kernels_match_oracle(B, cmax, case):
    """Edge cases of the SupCon reduction against the oracle: more rows than the CTA has warps, Cmax 300, multi-positive
    rows, rows without positives, and the batches the reference maps to zero_losses() (components/losses.py:14-15,20)."""
    from oracle import nrms_oracle as O
    from newsreclib_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + cmax)
    cnt = torch.randint(1, cmax + 1, (B,), generator=g) if case != "train" else torch.full((B,), cmax)
    cnt[0] = cmax
    seg = torch.repeat_interleave(torch.arange(B), cnt)
    labels = (torch.rand(int(cnt.sum()), generator=g) < 0.2).float()
    if case == "no_pos":
        labels.zero_()
    elif case == "no_neg":
        labels.fill_(1.0)
    elif case == "one_pair":                                               # every index list has one element: zero_losses()
        labels = torch.tensor([1.0, 0.0])
    mask = torch.arange(cmax)[None, :] < cnt[:, None]
    scores = torch.randn(B, cmax, generator=g) * 3 * mask                   # padded slots score exactly 0
    for coef in (None, 0.3):
        sr = scores.clone().requires_grad_(True)
        y, mk = O.to_dense_batch(labels, seg)
        ref = O.sup_con_loss(sr, y, mk) if coef is None else (1 - coef) * O.ce_soft(sr, y) + coef * O.sup_con_loss(sr, y, mk)
        ref.backward()
        sg = scores.cuda().requires_grad_(True)
        off = ops.segment_offsets(seg.cuda(), B)
        loss = ops.SupConFn.apply(sg, labels.cuda(), off, coef)
        loss.backward()
        if float(ref) == 0.0:
            assert float(loss) == 0.0 and bool((sg.grad == 0).all())
        else:
            assert rel_err(loss.detach(), ref.detach()) <= 1e-5, (case, coef)
            assert rel_err(sg.grad, sr.grad) <= 1e-4, (case, coef)


@pytest.mark.parametrize("split", ["test", "train"])
def test_device_collate_matches_reference_collate_golden(split):
    """DeviceCollate against the output of the reference's OWN DatasetCollate (tests/golden/collate_ref.npz, minted by
    oracle/make_collate_golden.py): every tensor of the RecommendationBatch bit for bit, dtype included."""
    from helpers import load_collate_golden
    from newsreclib_b200.data.components.device_collate import DeviceCollate, DeviceNewsTable
    news, splits, (lt, la) = load_collate_golden()
    samples, ref = splits[split]
    table = DeviceNewsTable.from_token_lists(
        news["nid"], news["tokenized_title"], news["category_class"], news["subcategory_class"], lt,
        news["tokenized_abstract"], la, news["sentiment_class"], news["sentiment_score"])
    got = DeviceCollate(table)(samples)
    for k, v in ref.items():
        if isinstance(v, dict):
            assert set(got[k]) == set(v)
            for c, t in v.items():
                assert got[k][c].dtype == t.dtype and torch.equal(got[k][c].cpu(), t), (k, c)
        else:
            assert got[k].dtype == v.dtype and torch.equal(got[k].cpu(), v), k
