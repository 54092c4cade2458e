"""GPU parity tests (-m gpu): the CUDA path through the C ABI against the CPU oracle and the
committed golden fixtures.  Tolerances (BASELINE.json north_star / SURVEY.md §8d):
  index/gather work bit-exact; fp32 logits max|s-s_ref|/max|s_ref| <= 1e-4; loss rel 1e-4;
  gradients rel 1e-3 (fp32-parity mode = three bf16 tensor-core passes on hi/lo split operands);
  bf16 single-pass mode is held to 2e-2 on logits.
"""
import numpy as np
import pytest
import torch

from helpers import TITLE, USER, batch_sizes, gpu_run, grad_tolerances, load_golden, oracle_run, rel_err, to_dev
from newsreclib_b200.synthetic import make_batch, make_nrms_params

pytestmark = pytest.mark.gpu

LOGIT_TOL, LOSS_TOL, GRAD_TOL, BF16_TOL = 1e-4, 1e-4, 1e-3, 2e-2


@pytest.mark.parametrize("shape", [(128, 16, 16), (130, 64, 64), (256, 160, 304), (300, 208, 304),
                                   (1000, 900, 304), (777, 300, 912), (5000, 240, 300)])
def test_gemm_nt_tcgen05(shape):
    from newsreclib_b200 import ops
    M, N, K = shape
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    ref = (A.double() @ B.double().t())
    D3 = ops.gemm_test(A, B, False, ops.PREC_BF16X3)
    D1 = ops.gemm_test(A, B, False, ops.PREC_BF16)
    torch.cuda.synchronize()
    e3, e1 = rel_err(D3, ref), rel_err(D1, ref)
    print(f"NT {shape}: bf16x3 rel {e3:.2e}  bf16 rel {e1:.2e}")
    assert e3 < 3e-5 and e1 < 2e-2


@pytest.mark.parametrize("shape", [(19000, 304, 304), (19137, 900, 304), (40000, 208, 208), (19000, 300, 912)])
def test_gemm_nt_cta_pair(shape):
    """M >= 74 row-pairs of 256 routes the NT GEMM to the cta_group::2 kernel (CTA pairs, M = 256 MMAs,
    half of the B tile per CTA); ragged M (19137) exercises the all-out-of-bounds peer tile."""
    from newsreclib_b200 import ops
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    ref = (A.double() @ B.double().t())
    D3 = ops.gemm_test(A, B, False, ops.PREC_BF16X3)
    D1 = ops.gemm_test(A, B, False, ops.PREC_BF16)
    torch.cuda.synchronize()
    assert rel_err(D3, ref) < 3e-5 and rel_err(D1, ref) < 2e-2
    out = ops.gemm_test_planes(A, B).cpu()
    assert rel_err(out[:, :N], ref) < 4e-5
    assert torch.all(out[:, N] == 1.0) and torch.all(out[:, N + 1:] == 0.0)


@pytest.mark.parametrize("shape", [(900, 304, 20000), (200, 304, 12000), (400, 912, 10048), (256, 100, 9500)])
def test_gemm_tn_cta_pair(shape):
    """Weight-gradient (MN-major, split-K) GEMMs whose M fills whole 256-row pairs run on the CTA-pair kernel:
    BN = 256 / 128 tiles with a narrower last tile (304 = 256 + 48), TMA reduce-add of the k-split partials."""
    from newsreclib_b200 import ops
    M, N, K = shape
    g = torch.Generator().manual_seed(M * 5 + N)
    A = torch.randn(K, M, generator=g).cuda()
    B = torch.randn(K, N, generator=g).cuda()
    ref = (A.double().t() @ B.double())
    D3 = ops.gemm_test(A, B, True, ops.PREC_BF16X3)
    D1 = ops.gemm_test(A, B, True, ops.PREC_BF16)
    torch.cuda.synchronize()
    assert rel_err(D3, ref) < 3e-5 and rel_err(D1, ref) < 2e-2


@pytest.mark.parametrize("shape", [(128, 16, 16), (300, 208, 304), (1000, 300, 304), (4000, 900, 304)])
def test_gemm_plane_sink(shape):
    """The bf16 hi/lo plane sink (TMA store of a [2, 32, 32] box): values, ones column, zero pad."""
    from newsreclib_b200 import ops
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    ref = (A.double() @ B.double().t())
    out = ops.gemm_test_planes(A, B).cpu()
    assert rel_err(out[:, :N], ref) < 4e-5          # 3-pass product, then rounded to 16 mantissa bits
    assert torch.all(out[:, N] == 1.0) and torch.all(out[:, N + 1:] == 0.0)
    out1 = ops.gemm_test_planes(A, B, ops.PREC_BF16).cpu()
    assert rel_err(out1[:, :N], ref) < 2e-2 and torch.all(out1[:, N] == 1.0)


@pytest.mark.parametrize("shape", [(128, 64, 64), (180, 64, 333), (200, 304, 777), (900, 304, 5000),
                                   (300, 304, 20000)])
def test_gemm_tn_tcgen05(shape):
    from newsreclib_b200 import ops
    M, N, K = shape
    g = torch.Generator().manual_seed(M * 3 + N)
    A = torch.randn(K, M, generator=g).cuda()
    B = torch.randn(K, N, generator=g).cuda()
    ref = (A.double().t() @ B.double())
    D3 = ops.gemm_test(A, B, True, ops.PREC_BF16X3)
    D1 = ops.gemm_test(A, B, True, ops.PREC_BF16)
    torch.cuda.synchronize()
    e3, e1 = rel_err(D3, ref), rel_err(D1, ref)
    print(f"TN {shape}: bf16x3 rel {e3:.2e}  bf16 rel {e1:.2e}")
    assert e3 < 3e-5 and e1 < 2e-2


def _compare(scores, loss, grads, ref_scores, ref_loss, ref_grads, tag, tols=None):
    es, el = rel_err(scores, ref_scores), rel_err(loss, ref_loss)
    print(f"{tag}: logits rel {es:.2e}  loss rel {el:.2e}")
    assert es <= LOGIT_TOL and el <= LOSS_TOL
    if grads is not None:
        worst = 0.0
        for k, g in ref_grads.items():
            e, t = rel_err(grads[k], g), (tols[k] if tols else GRAD_TOL)
            if t > GRAD_TOL:  # a bar relaxed to the fp32 oracle's own noise is always shown
                print(f"  grad {k}: err {e:.2e} tol {t:.2e} (4 x oracle-vs-fp64 noise)")
            worst = max(worst, e / t)
            assert e <= t, (tag, k, e, t)
        print(f"{tag}: worst gradient error / tolerance {worst:.2f}")


@pytest.mark.parametrize("name", ["nrms_tiny", "nrms_mind", "nrms_b8"])
def test_nrms_step_against_reference_golden(name):
    g, params, batch, d = load_golden(name)
    scores, loss, grads = gpu_run(params, batch, d["H"])
    es, el = rel_err(scores, g["scores"]), rel_err(loss, g["loss"])
    print(f"{name}: logits rel {es:.2e} loss rel {el:.2e}")
    assert es <= LOGIT_TOL and el <= LOSS_TOL
    B = d["B"]
    cnt = torch.bincount(batch["batch_cand"], minlength=B)
    for b in range(B):  # padded slots are exactly 0.0 (click_predictor.py:10 on zero rows)
        assert torch.all(scores[b, cnt[b]:] == 0)
    tols = grad_tolerances(params, batch, d["H"], GRAD_TOL)
    for k, v in g.items():
        if k.startswith("grad/"):
            assert rel_err(grads[k[5:]], v) <= tols[k[5:]], k
        elif k.startswith("gradsample/"):
            assert rel_err(grads[k[11:]].reshape(-1)[::7], v) <= tols[k[11:]], k
    assert float(grads[TITLE + "embedding_layer.weight"][0].abs().max()) == 0.0  # padding_idx row


@pytest.mark.parametrize("hist,cand,B", [("ragged", "train", 16), ("fixed", "train", 8), ("ragged", "eval", 6)])
def test_nrms_step_against_oracle(hist, cand, B):
    V = 3000
    params = make_nrms_params(V, seed=B)
    batch = make_batch(B, V, hist=hist, cand=cand, seed=100 + B, max_hist=20)
    rs, rl, rg = oracle_run(params, batch, 15)
    scores, loss, grads = gpu_run(params, batch, 15)
    _compare(scores, loss, grads, rs, rl, rg, f"oracle[{hist},{cand},B={B}]",
             grad_tolerances(params, batch, 15, GRAD_TOL, rg))


@pytest.mark.parametrize("B,max_hist,hist,cand,L", [(1, 1, "fixed", "train", 30), (2, 3, "ragged", "eval", 5),
                                                     (40, 4, "ragged", "train", 9), (3, 50, "fixed", "eval", 30)])
def test_nrms_step_edge_shapes(B, max_hist, hist, cand, L):
    """Edge cases: a single impression (the batch-axis user attention degenerates to S = 1), one-news
    histories, very short titles, B > 32 (tile attention kernel for the user encoder), long candidate lists."""
    V = 400
    params = make_nrms_params(V, seed=B * 7 + L)
    batch = make_batch(B, V, hist=hist, cand=cand, seed=B + L, max_hist=max_hist, max_title_len=L)
    scores, loss, grads = gpu_run(params, batch, 15)
    rs, rl, rg = oracle_run(params, batch, 15)
    _compare(scores, loss, grads, rs, rl, rg, f"edge B={B} hist<={max_hist} L={L}",
             grad_tolerances(params, batch, 15, GRAD_TOL, rg))


def test_nrms_step_bf16_mode():
    from newsreclib_b200 import ops
    V = 2000
    params = make_nrms_params(V, seed=9)
    batch = make_batch(8, V, hist="ragged", seed=9, max_hist=12)
    rs, rl, _ = oracle_run(params, batch, 15, grad=False)
    scores, loss, _ = gpu_run(params, batch, 15, precision=ops.PREC_BF16, do_backward=False)
    e = rel_err(scores, rs)
    print(f"bf16 single-pass: logits rel {e:.2e}")
    assert e <= BF16_TOL


def test_nrms_late_fusion():
    V = 1500
    params = make_nrms_params(V, seed=4)
    batch = make_batch(6, V, hist="ragged", seed=4, max_hist=9)
    rs, rl, rg = oracle_run(params, batch, 15, late_fusion=True)
    scores, loss, grads = gpu_run(params, batch, 15, late_fusion=True)
    rg = {k: v for k, v in rg.items() if k.startswith(TITLE)}
    _compare(scores, loss, grads, rs, rl, rg, "late_fusion")


def test_train_mode_dropout_same_mask():
    """Train-mode parity: feed the oracle the very keep-masks the kernels draw."""
    from newsreclib_b200 import ops
    V, p, seed = 1200, 0.2, 77
    params = make_nrms_params(V, seed=6)
    batch = make_batch(5, V, hist="ragged", seed=6, max_hist=7)
    nh, L = batch["x_hist"]["title"].shape
    nc = batch["x_cand"]["title"].shape[0]
    E = 300
    n = (nh + nc) * L * E  # history and candidate rows are encoded in one pass, history first
    m0 = ops.dropout_mask(n, seed, 0, p, "cuda").cpu().float().reshape(nh + nc, L, E)
    m1 = ops.dropout_mask(n, seed, 1, p, "cuda").cpu().float().reshape(nh + nc, L, E)
    frac = float(m0.mean())
    assert abs(frac - (1 - p)) < 5e-3, frac
    masks = {"hist1": m0[:nh], "hist2": m1[:nh], "cand1": m0[nh:], "cand2": m1[nh:]}
    rs, rl, rg = oracle_run(params, batch, 15, masks=masks, dropout_p=p)
    scores, loss, grads = gpu_run(params, batch, 15, dropout_p=p, training=True, seed=seed)
    _compare(scores, loss, grads, rs, rl, rg, "train-mode dropout")


def test_full_size_properties():
    """BASELINE config 2 size (B=64, hist 50, 5 candidates): size-independent properties."""
    from newsreclib_b200 import ops
    V = 70000
    params = make_nrms_params(V, seed=1234)
    batch = make_batch(64, V, hist="fixed", seed=1234)
    scores, loss, grads = gpu_run(params, batch, 15)
    assert torch.isfinite(scores).all() and torch.isfinite(loss)
    assert scores.shape == (64, 5)
    # (1) permutation equivariance over impressions: the batch-axis attention has no positional
    #     information, so permuting the impressions permutes the scores
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(0))
    nh = batch["x_hist"]["title"].shape[0]
    hist_rows = torch.arange(nh).reshape(64, 50)[perm].reshape(-1)
    cand_rows = torch.arange(320).reshape(64, 5)[perm].reshape(-1)
    b2 = {"x_hist": {"title": batch["x_hist"]["title"][hist_rows]},
          "x_cand": {"title": batch["x_cand"]["title"][cand_rows]},
          "batch_hist": batch["batch_hist"], "batch_cand": batch["batch_cand"],
          "labels": batch["labels"][cand_rows]}
    s2, l2, g2 = gpu_run(params, b2, 15)
    assert rel_err(s2, scores[perm]) < 2e-5
    assert rel_err(l2, loss) < 1e-5
    # (2) embedding-gradient rows of tokens absent from the batch, and row 0, are exactly zero
    gt = grads[TITLE + "embedding_layer.weight"]
    used = torch.zeros(V + 1, dtype=torch.bool)
    used[batch["x_hist"]["title"].reshape(-1)] = True
    used[batch["x_cand"]["title"].reshape(-1)] = True
    assert float(gt[~used].abs().max()) == 0.0 and float(gt[0].abs().max()) == 0.0
    assert rel_err(g2[TITLE + "embedding_layer.weight"], gt) < 1e-3
    # (3) sum over the batch of d loss / d scores is ~0 per row (softmax minus one-hot)
    # (4) loss equals the oracle CE of the returned scores
    from oracle import nrms_oracle as O
    assert rel_err(O.nrms_loss(batch, scores), loss) < 1e-5


def test_user_coupling_quirk_on_gpu(golden_dir):
    import os
    from newsreclib_b200 import ops
    g = dict(np.load(os.path.join(golden_dir, "user_coupling.npz")))
    E, H, Q = [int(x) for x in g["meta"]]
    blk = [torch.from_numpy(g["param/" + k]).cuda() for k in ops.BLOCK_KEYS]
    for key_in, key_out in (("h", "u"), ("h2", "u2")):
        u = ops.UserEncoderFn.apply(torch.from_numpy(g[key_in]).cuda(), *blk, H, 0, ops.PREC_BF16X3)
        assert rel_err(u, g[key_out]) <= LOGIT_TOL


def test_errors_are_loud():
    from newsreclib_b200 import ops
    with pytest.raises(RuntimeError):
        ops.nrms_step({"x_hist": {"title": torch.zeros(2, 3, dtype=torch.int64)}, "x_cand": {"title": torch.zeros(2, 3, dtype=torch.int64)},
                       "batch_hist": torch.zeros(2, dtype=torch.int64), "batch_cand": torch.zeros(2, dtype=torch.int64),
                       "labels": torch.zeros(2)}, torch.zeros(4, 60), [torch.zeros(1)] * 7, None, ops.dims_of(60, 3, 40),
                      B=1, Hmax=2, Cmax=2)
    with pytest.raises(RuntimeError, match="head dim"):
        ops.UserEncoderFn.apply(torch.zeros(2, 2, 70).cuda(), *[torch.zeros(4).cuda()] * 7, 2, 0, 0)
