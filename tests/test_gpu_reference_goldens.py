"""GPU tests against fixtures that hold the output of the reference's OWN top-level code, run in the build container
under stand-ins for its absent third-party imports: ``NRMSModule.forward`` / ``model_step``
(``oracle/make_module_golden.py``) and ``DatasetCollate`` (``oracle/make_collate_golden.py``)."""
import pytest
import torch

from helpers import grad_tolerances, oracle_run, rel_err
from test_gpu_modules import full_batch, make_module

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["nrms_module_ref", "nrms_module_ref_late_fusion"])
def test_nrms_module_matches_reference_module_golden(name):
    """The drop-in NRMSModule against what the reference's OWN NRMSModule.forward / model_step returned on the same
    batch and weights (tests/golden/nrms_module_ref*.npz, minted by oracle/make_module_golden.py): scores, loss, every
    entry of the 11-tuple, and the gradients."""
    from helpers import MODULE_OUT, USER, grad_sample, load_module_golden
    params, batch, ref, meta = load_module_golden(name)
    lf = meta["late_fusion"]
    use = {k: v for k, v in params.items() if not (lf and k.startswith(USER))}
    m = make_module(params, late_fusion=lf)
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
    _, _, rg = oracle_run(use, batch, meta["H"], late_fusion=lf)
    tols = grad_tolerances(use, batch, meta["H"], 1e-3, rg, late_fusion=lf)
    for k, p in m.named_parameters():
        g = ref["grad"][k]
        if float(g.abs().max()) < 1e-9:
            continue  # mathematically zero gradient (key bias): rounding noise on both sides
        assert p.grad is not None, k
        # 2 x: the stored sample of a big gradient normalises by the sample's maximum, not the tensor's
        assert rel_err(grad_sample(p.grad), g) <= 2.0 * tols[k], k


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
