"""GPU parity tests (-m gpu) at the sizes bench.py measures: BASELINE.json configs[1] (NRMS, MINDsmall-shape: B = 64,
history 50, 5 candidates, 30-token titles, V = 70 000) and configs[2] (MINDlarge-shape, V = 130 000, bf16), the CUDA path
through the C ABI against the CPU oracle on the same seeded inputs.  At these sizes every title-block GEMM runs on the
CTA-pair kernel (``nrl_gemm_tc2_kernel``) with its fused epilogues (dropout keep-bits, tanh.q score, rank-1 addend,
bias-gradient column, fused n-tiles), which the small fixtures never reach.

Tolerances (BASELINE.json north_star / SURVEY.md section 8d): gathers bit-exact; fp32 logits
max|s - s_ref| / max|s_ref| <= 1e-4; loss rel 1e-4; gradients rel 1e-3 -- relaxed per tensor to 4 x the fp32 oracle's
own distance from an fp64 run of the same oracle where that is larger (printed for every tensor); bf16 mode 2e-2
against the oracle under torch.autocast(bfloat16).
"""
import numpy as np
import pytest
import torch

from helpers import TITLE, USER, batch_sizes, gpu_run, grad_tolerances, oracle_run, rel_err, report_step
from newsreclib_b200.synthetic import make_batch, make_nrms_params, make_titles

pytestmark = pytest.mark.gpu

LOGIT_TOL, LOSS_TOL, GRAD_TOL, BF16_TOL = 1e-4, 1e-4, 1e-3, 2e-2
V_SMALL, V_LARGE = 70000, 130000


# ------------------------------------------------------------------------------------------------ a2: the gather
def test_embedding_gather_bit_exact():
    """nn.Embedding.forward (text.py:215-217,224): the kernel's fp32 sink is table[ids] bit for bit (row 0 is a real
    row), and the bf16 hi / lo planes the GEMMs consume equal the host-side split bit for bit."""
    from newsreclib_b200 import ops
    g = torch.Generator().manual_seed(5)
    V, E = V_SMALL, 300
    table = torch.randn(V + 1, E, generator=g)
    ids = torch.randint(0, V + 1, (4096, 30), generator=g)
    ids[0, :4] = torch.tensor([0, V, 0, 1])  # the padding row, the last row
    ids[:, 20:] = 0                          # right padding, as the collate emits
    out, hi, lo = ops.embedding_gather(ids.cuda(), table.cuda(), want_planes=True)
    torch.cuda.synchronize()
    ref = table[ids.reshape(-1)]
    assert torch.equal(out.cpu(), ref)
    ref_hi = ref.to(torch.bfloat16)
    ref_lo = (ref - ref_hi.float()).to(torch.bfloat16)
    hi, lo = hi.cpu(), lo.cpu()
    assert torch.equal(hi[:, :E].view(torch.int16), ref_hi.view(torch.int16))
    assert torch.equal(lo[:, :E].view(torch.int16), ref_lo.view(torch.int16))
    assert torch.all(hi[:, E].float() == 1.0) and torch.all(hi[:, E + 1:].float() == 0.0)  # bias column, zero pad
    assert torch.all(lo[:, E:].float() == 0.0)
    assert ops.device_status(raise_on_error=False) == 0


def test_device_side_input_checks():
    """nn.Embedding raises on an id outside the table; here the kernels never read or write out of bounds, set a sticky
    device word, and the host call reports it (ids were never range-checked)."""
    from newsreclib_b200 import ops
    V = 500
    params = make_nrms_params(V, seed=2)
    batch = make_batch(4, V, hist="ragged", seed=2, max_hist=6)
    assert ops.device_status(raise_on_error=False) == 0
    bad = {k: (dict(v) if isinstance(v, dict) else v) for k, v in batch.items()}
    bad["x_hist"]["title"] = batch["x_hist"]["title"].clone()
    bad["x_hist"]["title"][1, 2] = V + 7
    s_bad, _, g_bad = gpu_run(params, bad, 15)
    assert torch.isfinite(s_bad).all()
    with pytest.raises(RuntimeError, match="token id"):
        ops.device_status()
    assert ops.device_status(raise_on_error=False) == 0          # reading clears it
    bad2 = dict(batch)
    bad2["batch_cand"] = batch["batch_cand"].clone()
    bad2["batch_cand"][:] = 0                                     # one impression with every candidate: longer than Cmax
    B, Hmax, Cmax = batch_sizes(batch)
    P = {k: v.cuda() for k, v in params.items()}
    from helpers import to_dev
    ops.nrms_step(to_dev(bad2), P[TITLE + "embedding_layer.weight"], ops.block_from_dict(P, TITLE),
                  ops.block_from_dict(P, USER), ops.dims_of(300, 15, 200), B=B, Hmax=Hmax, Cmax=Cmax)
    assert ops.device_status(raise_on_error=False) in (2, 3)
    # a well-formed step afterwards is clean
    gpu_run(params, batch, 15)
    assert ops.device_status(raise_on_error=False) == 0


# ------------------------------------------------------------------------------- configs[1] at the measured size
@pytest.mark.parametrize("hist", ["fixed", "ragged"])
def test_benchmark_size_eval_mode_vs_oracle(hist):
    params = make_nrms_params(V_SMALL, seed=1234)
    batch = make_batch(64, V_SMALL, hist=hist, cand="train", seed=1234)
    assert batch["x_hist"]["title"].shape[0] * 30 >= 2 * 74 * 256 or hist == "ragged"  # CTA-pair GEMMs (>= 74 row pairs)
    rs, rl, rg = oracle_run(params, batch, 15)
    scores, loss, grads = gpu_run(params, batch, 15)
    tols = grad_tolerances(params, batch, 15, GRAD_TOL, rg)
    report_step(f"B=64 hist={hist} V={V_SMALL} eval", scores, loss, grads, rs, rl, rg, tols, LOGIT_TOL, LOSS_TOL)
    assert float(grads[TITLE + "embedding_layer.weight"][0].abs().max()) == 0.0


def test_benchmark_size_train_mode_same_masks():
    """The benchmarked step itself (train mode, dropout 0.2): the oracle is fed the very keep-masks the kernels draw."""
    from newsreclib_b200 import ops
    p, seed = 0.2, 20261017
    params = make_nrms_params(V_SMALL, seed=1234)
    batch = make_batch(64, V_SMALL, hist="fixed", cand="train", seed=4321)
    nh, L = batch["x_hist"]["title"].shape
    nc = batch["x_cand"]["title"].shape[0]
    E = 300
    n = (nh + nc) * L * E
    m0 = ops.dropout_mask(n, seed, 0, p, "cuda").cpu().float().reshape(nh + nc, L, E)
    m1 = ops.dropout_mask(n, seed, 1, p, "cuda").cpu().float().reshape(nh + nc, L, E)
    masks = {"hist1": m0[:nh], "hist2": m1[:nh], "cand1": m0[nh:], "cand2": m1[nh:]}
    rs, rl, rg = oracle_run(params, batch, 15, masks=masks, dropout_p=p)
    scores, loss, grads = gpu_run(params, batch, 15, dropout_p=p, training=True, seed=seed)
    tols = grad_tolerances(params, batch, 15, GRAD_TOL, rg, masks=masks, dropout_p=p)
    report_step("B=64 hist=fixed train-mode p=0.2", scores, loss, grads, rs, rl, rg, tols, LOGIT_TOL, LOSS_TOL)


def _eval_batch_with_long_impression(B, V, cmax, seed):
    """Evaluation-style batch (whole impressions as candidates) whose last impression has `cmax` candidates."""
    batch = make_batch(B, V, hist="ragged", cand="eval", seed=seed)
    cnt = torch.bincount(batch["batch_cand"], minlength=B)
    extra = cmax - int(cnt[-1])
    assert extra > 0
    rng = np.random.default_rng(seed + 1)
    out = {k: (dict(v) if isinstance(v, dict) else v) for k, v in batch.items() if k != "dense_widths"}  # widths change below
    out["x_cand"] = {"title": torch.cat([batch["x_cand"]["title"], torch.from_numpy(make_titles(rng, extra, V))])}
    out["x_hist"] = {"title": batch["x_hist"]["title"]}
    out["labels"] = torch.cat([batch["labels"], torch.zeros(extra)])
    out["batch_cand"] = torch.cat([batch["batch_cand"], torch.full((extra,), B - 1, dtype=torch.int64)])
    return out


def test_eval_batch_cmax_300_vs_oracle():
    """Evaluation-shaped batch (Cmax up to ~300 on MINDlarge-dev): dense [B, 300] scores with exact
    zeros in the padded slots, CE over the padded row, and the backward through the wide scorer."""
    B = 64
    params = make_nrms_params(V_SMALL, seed=99)
    batch = _eval_batch_with_long_impression(B, V_SMALL, 300, seed=99)
    _, Hmax, Cmax = batch_sizes(batch)
    assert Cmax == 300
    rs, rl, rg = oracle_run(params, batch, 15)
    scores, loss, grads = gpu_run(params, batch, 15)
    cnt = torch.bincount(batch["batch_cand"], minlength=B)
    for b in range(B):
        assert torch.all(scores[b, cnt[b]:] == 0)
    tols = grad_tolerances(params, batch, 15, GRAD_TOL, rg)
    report_step(f"eval batch B=64 Cmax=300 Hmax={Hmax}", scores, loss, grads, rs, rl, rg, tols, LOGIT_TOL, LOSS_TOL)


# ------------------------------------------------------------------------------------- configs[2]: bf16 single pass
def _autocast_oracle(params, batch, H):
    """The oracle under torch.autocast(bfloat16) (what the reference's precision-16 trainers run)."""
    from oracle import nrms_oracle as O
    ps = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    with torch.autocast("cpu", dtype=torch.bfloat16):
        scores = O.nrms_forward(batch, ps, H)
        loss = O.nrms_loss(batch, scores.float())
    loss.backward()
    grads = {k: (v.grad.detach().float() if v.grad is not None else torch.zeros_like(v)) for k, v in ps.items()}
    grads[TITLE + "embedding_layer.weight"][0] = 0
    return scores.detach().float(), loss.detach().float(), grads


def test_bf16_mode_forward_backward_vs_autocast_oracle():
    """NRL_PREC_BF16 (one bf16 plane, one MMA per k-step) at MINDlarge-shape, forward AND backward, against the oracle
    under torch.autocast(bfloat16) at 2e-2.  Two bf16 evaluations of the same function round differently, so a
    gradient's bar is never tighter than twice the autocast oracle's own distance from the fp32 oracle (printed)."""
    from newsreclib_b200 import ops
    params = make_nrms_params(V_LARGE, seed=7)
    batch = make_batch(64, V_LARGE, hist="ragged", cand="train", seed=7)
    as_, al, ag = _autocast_oracle(params, batch, 15)
    fs, fl, fg = oracle_run(params, batch, 15)
    scores, loss, grads = gpu_run(params, batch, 15, precision=ops.PREC_BF16)
    e_auto, e_f32, own = rel_err(scores, as_), rel_err(scores, fs), rel_err(as_, fs)
    print(f"bf16 logits: vs autocast oracle {e_auto:.2e}, vs fp32 oracle {e_f32:.2e} (autocast oracle vs fp32 oracle {own:.2e})")
    assert e_auto <= BF16_TOL and e_f32 <= BF16_TOL
    assert rel_err(loss, al) <= BF16_TOL and rel_err(loss, fl) <= BF16_TOL
    worst = 0.0
    for k, g in ag.items():
        if float(fg[k].abs().max()) < 1e-9:
            continue  # mathematically zero (key bias)
        own = rel_err(g, fg[k])
        tol = max(BF16_TOL, 2.0 * own)
        e = min(rel_err(grads[k], g), rel_err(grads[k], fg[k]))
        print(f"  bf16 grad {k:<62s} err {e:.2e}  tol {tol:.2e}  (autocast oracle vs fp32 oracle {own:.2e})")
        assert e <= tol, (k, e, tol)
        worst = max(worst, e / tol)
    print(f"bf16 fwd+bwd: worst gradient error / tolerance {worst:.2f}")
    assert float(grads[TITLE + "embedding_layer.weight"][0].abs().max()) == 0.0
