"""Torch-facing wrappers of the C ABI: raw calls on ``torch`` CUDA tensors (device pointers
and the current stream are handed to the library; torch only owns the memory) and the
``torch.autograd.Function``s the drop-in modules use.

No function here computes anything on the CPU or through ATen kernels on the hot path; a
missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import BlockParams, CnnDims, CnnParams, Dims, PREC_BF16, PREC_BF16X3  # noqa: F401

BLOCK_KEYS = (
    "multihead_attention.in_proj_weight",
    "multihead_attention.in_proj_bias",
    "multihead_attention.out_proj.weight",
    "multihead_attention.out_proj.bias",
    "additive_attention.linear.weight",
    "additive_attention.linear.bias",
    "additive_attention.query",
)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (newsreclib_b200 has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _ws(ctx) -> torch.Tensor:
    """The workspace a Function's forward left on ``ctx`` (it holds the saved activations and is released after the first
    backward, so a second backward through the same node -- ``retain_graph=True`` -- cannot be served)."""
    if ctx.ws is None:
        raise RuntimeError("backward called twice through a newsreclib_b200 op: the saved activations live in a workspace "
                           "that is released after the first backward (retain_graph is not supported)")
    return ctx.ws


def block_struct(tensors) -> BlockParams:
    """Seven tensors in ``BLOCK_KEYS`` order -> ``nrl_block_params``."""
    s = BlockParams()
    for (name, _), t in zip(BlockParams._fields_, tensors):
        setattr(s, name, _p(_chk(t, torch.float32, name)))
    return s


def block_from_dict(params: Dict[str, torch.Tensor], prefix: str):
    return [params[prefix + k] for k in BLOCK_KEYS]


def workspace(nbytes: int, device) -> torch.Tensor:
    # torch's caching allocator returns 512-byte aligned blocks; over-allocate and slice to 1 KiB
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    shift = (-buf.data_ptr()) % 1024
    return buf[shift: shift + nbytes]


def dims_of(E: int, H: int, Q: int) -> Dims:
    return Dims(int(E), int(H), int(Q))


# ----------------------------------------------------------------------------------------
# raw calls
# ----------------------------------------------------------------------------------------
def gemm_test(A: torch.Tensor, B: torch.Tensor, mn_major: bool, precision: int = PREC_BF16X3) -> torch.Tensor:
    lib = _lib.load()
    _chk(A, torch.float32, "A"); _chk(B, torch.float32, "B")
    if not mn_major:
        M, K = A.shape; N = B.shape[0]
    else:
        K, M = A.shape; N = B.shape[1]
    D = torch.zeros(M, N, dtype=torch.float32, device=A.device)
    ws = workspace(lib.nrl_gemm_test_ws_bytes(M, N, K), A.device)
    _lib.check(lib.nrl_gemm_test(_p(A), _p(B), _p(D), M, N, K, int(mn_major), precision, _p(ws),
                                 ws.numel(), _stream()), "nrl_gemm_test")
    return D


def gemm_test_planes(A: torch.Tensor, B: torch.Tensor, precision: int = PREC_BF16X3) -> torch.Tensor:
    """NT GEMM through the split-plane sink; returns hi + lo as fp32 ``[M, round_up(N + 1, 16)]``."""
    lib = _lib.load()
    _chk(A, torch.float32, "A"); _chk(B, torch.float32, "B")
    M, K = A.shape
    N = B.shape[0]
    Np = (N + 1 + 15) // 16 * 16
    out = torch.empty(M, Np, dtype=torch.float32, device=A.device)
    ws = workspace(lib.nrl_gemm_test_ws_bytes(M, N, K), A.device)
    _lib.check(lib.nrl_gemm_test_planes(_p(A), _p(B), _p(out), M, N, K, precision, _p(ws), ws.numel(),
                                        _stream()), "nrl_gemm_test_planes")
    return out


def dropout_mask(n: int, seed: int, site: int, p: float, device) -> torch.Tensor:
    lib = _lib.load()
    keep = torch.empty(n, dtype=torch.uint8, device=device)
    _lib.check(lib.nrl_dropout_mask(_p(keep), n, seed, site, p, _stream()), "nrl_dropout_mask")
    return keep


def segment_offsets(seg: torch.Tensor, B: int) -> torch.Tensor:
    lib = _lib.load()
    _chk(seg, torch.int64, "segment ids")
    off = torch.empty(B + 1, dtype=torch.int32, device=seg.device)
    _lib.check(lib.nrl_segment_offsets(_p(seg), seg.numel(), B, _p(off), _stream()), "nrl_segment_offsets")
    return off


def gather_rows(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """``table[idx]`` along dim 0 (``nrl_gather_rows``): token-id / scalar / news-vector rows."""
    lib = _lib.load()
    if not table.is_cuda or not table.is_contiguous():
        raise RuntimeError("table must be a contiguous CUDA tensor (newsreclib_b200 has no CPU path)")
    idx = _chk(idx.contiguous(), torch.int64, "idx")
    row_bytes = table[0].numel() * table.element_size() if table.dim() > 1 else table.element_size()
    if row_bytes % 4:
        raise RuntimeError("row size must be a multiple of 4 bytes")
    out = torch.empty((idx.numel(),) + tuple(table.shape[1:]), dtype=table.dtype, device=table.device)
    if idx.numel():
        _lib.check(lib.nrl_gather_rows(_p(table), table.shape[0], row_bytes, _p(idx), idx.numel(), _p(out), _stream()),
                   "nrl_gather_rows")
    return out


_weights_epoch = 0


def weights_epoch() -> int:
    """Counts the parameter updates this library made through raw pointers (``nrl_adam_step``,
    ``nrl_exchange_adam_step``): torch's version counters do not see those, so anything that caches a function of the
    weights (``TfmState``'s packed GEMM operands) keys on this as well."""
    return _weights_epoch


def bump_weights_epoch() -> None:
    global _weights_epoch
    _weights_epoch += 1


def adam_step(p, g, m, v, step: int, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0,
              zero_grad: bool = False) -> None:
    """``torch.optim.Adam.step`` over flat buffers; ``zero_grad=True`` also clears ``g`` in the same pass."""
    lib = _lib.load()
    bump_weights_epoch()
    for t, n in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _chk(t, torch.float32, n)
    fn = lib.nrl_adam_step_zero_grad if zero_grad else lib.nrl_adam_step
    _lib.check(fn(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, step, grad_scale, _stream()),
               "nrl_adam_step")


def embedding_gather(ids: torch.Tensor, table: torch.Tensor, want_planes: bool = False):
    """``nn.Embedding.forward`` (``text.py:224``): ``table[ids]`` as fp32 ``[n, E]`` and, optionally, the bf16 hi / lo
    split planes ``[n, Ep]`` the in-projection GEMM consumes (``nrl_embedding_gather``)."""
    lib = _lib.load()
    ids = _chk(ids.contiguous(), torch.int64, "ids")
    _chk(table, torch.float32, "embedding table")
    n, E = ids.numel(), table.shape[1]
    Ep = (E + 1 + 15) // 16 * 16
    out = torch.empty(n, E, dtype=torch.float32, device=table.device)
    hi = lo = None
    if want_planes:
        hi = torch.empty(n, Ep, dtype=torch.bfloat16, device=table.device)
        lo = torch.empty(n, Ep, dtype=torch.bfloat16, device=table.device)
    _lib.check(lib.nrl_embedding_gather(_p(ids), n, _p(table), table.shape[0], E, _p(out), _p(hi), _p(lo), _stream()),
               "nrl_embedding_gather")
    return (out, hi, lo) if want_planes else out


def device_status(raise_on_error: bool = True) -> int:
    """Synchronise the current stream and return (and clear) the device-side input-check word: 0, or the code of the
    first violation (token id outside the table, malformed segment ids ...) -- ``nrl_device_status``."""
    lib = _lib.load()
    code = C.c_int(0)
    _lib.check(lib.nrl_device_status(C.byref(code), _stream()), "nrl_device_status")
    if code.value and raise_on_error:
        raise RuntimeError("newsreclib_b200: " + lib.nrl_last_error().decode("utf-8", "replace"))
    return int(code.value)


def nrms_step(batch: Dict, table: torch.Tensor, news_block, user_block, dims: Dims, *, B: int,
              Hmax: int, Cmax: int, late_fusion: bool = False, dropout_p: float = 0.0,
              training: bool = False, seed: int = 0, want_loss: bool = True,
              grads: Optional[Tuple] = None, ws: Optional[torch.Tensor] = None,
              precision: int = PREC_BF16X3, scores: Optional[torch.Tensor] = None,
              loss: Optional[torch.Tensor] = None):
    """One pass of the hot path on device-resident inputs (``nrl_nrms_step``).

    ``grads`` = (news_grad_tensors[7], user_grad_tensors[7] or None, d_table) enables backward
    (gradients accumulate).  Returns (scores [B, Cmax], loss [1] or None, ws)."""
    lib = _lib.load()
    hist_ids = _chk(batch["x_hist"]["title"], torch.int64, "x_hist.title")
    cand_ids = _chk(batch["x_cand"]["title"], torch.int64, "x_cand.title")
    seg_h = _chk(batch["batch_hist"], torch.int64, "batch_hist")
    seg_c = _chk(batch["batch_cand"], torch.int64, "batch_cand")
    labels = _chk(batch["labels"], torch.float32, "labels")
    _chk(table, torch.float32, "table")
    nh, L = hist_ids.shape
    nc = cand_ids.shape[0]
    dev = table.device
    need = lib.nrl_nrms_ws_bytes(nh, nc, L, B, Hmax, Cmax, dims)
    if ws is None or ws.numel() < need:
        ws = workspace(need, dev)
    if scores is None:
        scores = torch.empty(B, Cmax, dtype=torch.float32, device=dev)
    if loss is None and (want_loss or grads is not None):
        loss = torch.empty(1, dtype=torch.float32, device=dev)
    nb = block_struct(news_block)
    ub = block_struct(user_block) if user_block is not None else None
    ng = ug = None
    d_table = None
    if grads is not None:
        ng = block_struct(grads[0])
        ug = block_struct(grads[1]) if grads[1] is not None else None
        d_table = grads[2]
    _lib.check(lib.nrl_nrms_step(
        _p(hist_ids), _p(cand_ids), _p(seg_h), _p(seg_c), _p(labels), nh, nc, L, B, Hmax, Cmax,
        _p(table), table.shape[0], C.byref(nb), C.byref(ub) if ub is not None else None, dims,
        int(late_fusion), float(dropout_p), int(training), int(seed), _p(scores), _p(loss),
        int(grads is not None), C.byref(ng) if ng is not None else None,
        C.byref(ug) if ug is not None else None, _p(d_table), _p(ws), ws.numel(), precision,
        _stream()), "nrl_nrms_step")
    return scores, loss, ws


# ----------------------------------------------------------------------------------------
# autograd functions used by the drop-in modules
# ----------------------------------------------------------------------------------------
class NrmsStepFn(torch.autograd.Function):
    """The differentiable part of ``NRMSModule.model_step`` (``nrms_module.py:230-255,277,288``: both news-encoder calls,
    ragged->dense, user encoder or late fusion, scorer, soft-target CE) as ONE autograd node on the fused C calls:
    ``forward`` = ``nrl_nrms_step`` without gradients (scores + loss; the workspace keeps every saved activation),
    ``backward`` = ``nrl_nrms_step_bwd`` with ``d objective / d loss`` read on the device.  Same kernels as the fused
    trainer step, so what a Lightning user drives (``training_step`` -> ``loss.backward()``) is the measured path and the
    host issues two library calls per step instead of ten ``autograd.Function``s.

    ``scores`` is returned for the metrics and is NOT differentiable here (a loss built from it needs the per-op path).
    ``grad_targets``: ``None`` -> gradients are returned to autograd (fresh zeroed buffers; what DDP / any optimizer
    expects); a list of 15 preallocated fp32 tensors (table, 7 title-block, 7 user-block or ``None``) -> the kernels
    accumulate (``+=``) straight into them and autograd gets ``None`` (``ModuleTrainer``: its flat gradient buffer)."""

    @staticmethod
    def forward(ctx, table, nw_in, nb_in, nw_out, nb_out, nw_add, nb_add, nq, uw_in, ub_in, uw_out, ub_out, uw_add,
                ub_add, uq, hist_ids, cand_ids, seg_h, seg_c, labels, cfg):
        B, Hmax, Cmax, num_heads, late_fusion, dropout_p, training, seed, precision, grad_targets = cfg
        news = [nw_in, nb_in, nw_out, nb_out, nw_add, nb_add, nq]
        user = None if late_fusion else [uw_in, ub_in, uw_out, ub_out, uw_add, ub_add, uq]
        dims = dims_of(table.shape[1], num_heads, nq.numel())
        batch = {"x_hist": {"title": hist_ids.contiguous()}, "x_cand": {"title": cand_ids.contiguous()},
                 "batch_hist": seg_h.contiguous(), "batch_cand": seg_c.contiguous(),
                 "labels": labels.float().contiguous()}
        scores, loss, ws = nrms_step(batch, table, news, user, dims, B=B, Hmax=Hmax, Cmax=Cmax, late_fusion=late_fusion,
                                     dropout_p=dropout_p, training=training, seed=seed, precision=precision)
        ctx.save_for_backward(table, *news, *(user or []), batch["labels"])
        ctx.ws, ctx.dims, ctx.cfg = ws, dims, cfg
        ctx.sizes = (hist_ids.shape[0], cand_ids.shape[0], hist_ids.shape[1])
        ctx.mark_non_differentiable(scores)
        return scores, loss.reshape(())

    @staticmethod
    def backward(ctx, _g_scores, g_loss):
        lib = _lib.load()
        B, Hmax, Cmax, num_heads, late_fusion, dropout_p, training, seed, precision, grad_targets = ctx.cfg
        if ctx.ws is None:
            raise RuntimeError("NrmsStepFn: backward called twice (the saved activations live in a workspace that is "
                               "released after the first backward; retain_graph is not supported on the fused step)")
        saved = ctx.saved_tensors
        table, news, labels = saved[0], list(saved[1:8]), saved[-1]
        user = None if late_fusion else list(saved[8:15])
        if grad_targets is None:
            d_table = torch.zeros_like(table)
            ng, ug = [torch.zeros_like(t) for t in news], (None if user is None else [torch.zeros_like(t) for t in user])
        else:
            d_table, ng, ug = grad_targets[0], list(grad_targets[1:8]), (None if user is None else list(grad_targets[8:15]))
            for t, p in zip([d_table] + ng + (ug or []), [table] + news + (user or [])):
                if t is None or t.shape != p.shape or t.dtype != torch.float32 or not t.is_contiguous() or t.device != p.device:
                    raise RuntimeError("NrmsStepFn: grad_targets must be contiguous fp32 tensors shaped like the parameters")
        g = g_loss.detach().float().reshape(1).contiguous()
        nh, nc, L = ctx.sizes
        nb, ub = block_struct(news), (block_struct(user) if user is not None else None)
        ngs, ugs = block_struct(ng), (block_struct(ug) if ug is not None else None)
        _lib.check(lib.nrl_nrms_step_bwd(
            _p(labels), _p(g), nh, nc, L, B, Hmax, Cmax, table.shape[0], C.byref(nb),
            C.byref(ub) if ub is not None else None, ctx.dims, int(late_fusion), float(dropout_p), int(training),
            int(seed), C.byref(ngs), C.byref(ugs) if ugs is not None else None, _p(d_table), _p(_ws(ctx)),
            ctx.ws.numel(), precision, _stream()), "nrl_nrms_step_bwd")
        ctx.ws = None
        if grad_targets is not None:
            return (None,) * 21
        return (d_table, *ng, *(ug if ug is not None else [None] * 7), None, None, None, None, None, None)


class NewsEncoderFn(torch.autograd.Function):
    """``MHSAAddAtt.forward`` (reference ``encoders/news/text.py:222-236``)."""

    @staticmethod
    def forward(ctx, ids, table, w_in, b_in, w_out, b_out, w_add, b_add, q_add, num_heads,
                dropout_p, training, seed, precision):
        lib = _lib.load()
        _chk(ids, torch.int64, "ids"); _chk(table, torch.float32, "embedding table")
        n, L = ids.shape
        dims = dims_of(table.shape[1], num_heads, q_add.numel())
        blk = [w_in, b_in, w_out, b_out, w_add, b_add, q_add]
        out = torch.empty(n, table.shape[1], dtype=torch.float32, device=table.device)
        ws = workspace(lib.nrl_news_encoder_ws_bytes(n, L, dims), table.device)
        bs = block_struct(blk)
        _lib.check(lib.nrl_news_encoder_fwd(_p(ids), n, L, _p(table), table.shape[0], C.byref(bs), dims,
                                            float(dropout_p), int(training), int(seed), _p(out), _p(ws),
                                            ws.numel(), precision, _stream()), "nrl_news_encoder_fwd")
        ctx.save_for_backward(ids, table, *blk)
        ctx.ws, ctx.dims, ctx.cfg = ws, dims, (float(dropout_p), int(training), int(seed), precision)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        ids, table, *blk = ctx.saved_tensors
        n, L = ids.shape
        p, training, seed, precision = ctx.cfg
        d_out = d_out.contiguous().float()
        grads = [torch.zeros_like(t) for t in blk]
        d_table = torch.zeros_like(table)
        bs, gs = block_struct(blk), block_struct(grads)
        _lib.check(lib.nrl_news_encoder_bwd(_p(ids), n, L, table.shape[0], C.byref(bs), ctx.dims, p, training,
                                            seed, _p(d_out), C.byref(gs), _p(d_table), _p(_ws(ctx)),
                                            ctx.ws.numel(), precision, _stream()), "nrl_news_encoder_bwd")
        ctx.ws = None
        return (None, d_table, *grads, None, None, None, None, None)


class UserEncoderFn(torch.autograd.Function):
    """NRMS ``UserEncoder.forward`` (reference ``encoders/user/nrms.py:32-41``)."""

    @staticmethod
    def forward(ctx, hist, w_in, b_in, w_out, b_out, w_add, b_add, q_add, num_heads, attention_axis,
                precision):
        lib = _lib.load()
        hist = _chk(hist.contiguous(), torch.float32, "hist")
        B, Hmax, E = hist.shape
        dims = dims_of(E, num_heads, q_add.numel())
        blk = [w_in, b_in, w_out, b_out, w_add, b_add, q_add]
        user = torch.empty(B, E, dtype=torch.float32, device=hist.device)
        ws = workspace(lib.nrl_user_encoder_ws_bytes(B, Hmax, dims), hist.device)
        bs = block_struct(blk)
        _lib.check(lib.nrl_user_encoder_fwd(_p(hist), B, Hmax, C.byref(bs), dims, int(attention_axis),
                                            _p(user), _p(ws), ws.numel(), precision, _stream()),
                   "nrl_user_encoder_fwd")
        ctx.save_for_backward(*blk)
        ctx.ws, ctx.dims, ctx.cfg = ws, dims, (B, Hmax, E, int(attention_axis), precision)
        return user

    @staticmethod
    def backward(ctx, d_user):
        lib = _lib.load()
        blk = list(ctx.saved_tensors)
        B, Hmax, E, axis, precision = ctx.cfg
        d_user = d_user.contiguous().float()
        grads = [torch.zeros_like(t) for t in blk]
        d_hist = torch.empty(B, Hmax, E, dtype=torch.float32, device=d_user.device)
        bs, gs = block_struct(blk), block_struct(grads)
        _lib.check(lib.nrl_user_encoder_bwd(B, Hmax, C.byref(bs), ctx.dims, axis, _p(d_user), C.byref(gs),
                                            _p(d_hist), _p(_ws(ctx)), ctx.ws.numel(), precision, _stream()),
                   "nrl_user_encoder_bwd")
        ctx.ws = None
        return (d_hist, *grads, None, None, None)


class ToDenseFn(torch.autograd.Function):
    """``to_dense_batch`` values (torch_geometric 2.3.0; call sites ``nrms_module.py:233,237``)."""

    @staticmethod
    def forward(ctx, x, off, B, M):
        lib = _lib.load()
        x = _chk(x.contiguous(), torch.float32, "x")
        E = x.shape[1]
        dense = torch.empty(B, M, E, dtype=torch.float32, device=x.device)
        _lib.check(lib.nrl_to_dense_fwd(_p(x), _p(off), B, M, E, _p(dense), _stream()), "nrl_to_dense_fwd")
        ctx.save_for_backward(off)
        ctx.cfg = (B, M, E, x.shape[0])
        return dense

    @staticmethod
    def backward(ctx, d_dense):
        lib = _lib.load()
        (off,) = ctx.saved_tensors
        B, M, E, n = ctx.cfg
        d_dense = d_dense.contiguous().float()
        dx = torch.zeros(n, E, dtype=torch.float32, device=d_dense.device)
        _lib.check(lib.nrl_to_dense_bwd(_p(d_dense), _p(off), B, M, E, _p(dx), _stream()), "nrl_to_dense_bwd")
        return dx, None, None, None


def rank_metrics(preds: torch.Tensor, targets: torch.Tensor, sizes: torch.Tensor, top_k_list, want_ranks: bool = False):
    """Per-impression reciprocal rank and nDCG@k (``nrl_rank_metrics``): ``[B, 1 + len(top_k_list)]`` (and, optionally,
    every candidate's rank within its impression)."""
    lib = _lib.load()
    preds = _chk(preds.detach().contiguous().float(), torch.float32, "preds")
    targets = targets.detach().to(preds.device).contiguous().float()
    sizes = sizes.to(preds.device).long()
    B = sizes.numel()
    off = torch.zeros(B + 1, dtype=torch.int64, device=preds.device)
    off[1:] = sizes.cumsum(0)
    ks = (C.c_int * max(1, len(top_k_list)))(*[int(k) for k in top_k_list])
    out = torch.empty(B, 1 + len(top_k_list), dtype=torch.float32, device=preds.device)
    ranks = torch.empty(preds.numel(), dtype=torch.int32, device=preds.device) if want_ranks else None
    _lib.check(lib.nrl_rank_metrics(_p(preds), _p(targets), _p(off), B, ks, len(top_k_list), _p(out), _p(ranks),
                                    _stream()), "nrl_rank_metrics")
    return (out, ranks) if want_ranks else out


def dense_to_ragged(dense: torch.Tensor, off: torch.Tensor, n: int) -> torch.Tensor:
    """Rows of a dense ``[B, M]`` / ``[B, M, E]`` batch back in ragged order (``[n]`` / ``[n, E]``): what boolean-mask
    indexing ``dense[mask]`` returns (``abstract_recommender.py:126-130``), without the host sync of ``nonzero`` -- the
    number of rows is known from the batch.  No gradient (used for the metric outputs)."""
    lib = _lib.load()
    dense = _chk(dense.detach().contiguous(), torch.float32, "dense")
    B, M = dense.shape[0], dense.shape[1]
    E = 1 if dense.dim() == 2 else dense.shape[2]
    out = torch.zeros((n,) if dense.dim() == 2 else (n, E), dtype=torch.float32, device=dense.device)
    _lib.check(lib.nrl_to_dense_bwd(_p(dense), _p(off), B, M, E, _p(out), _stream()), "nrl_to_dense_bwd")
    return out


class LateFusionFn(torch.autograd.Function):
    """Late fusion (``nrms_module.py:243-248``): user = mean of the impression's clicked-news vectors."""

    @staticmethod
    def forward(ctx, hist_vec, off, B):
        lib = _lib.load()
        hist_vec = _chk(hist_vec.contiguous(), torch.float32, "hist_vec")
        E = hist_vec.shape[1]
        user = torch.empty(B, E, dtype=torch.float32, device=hist_vec.device)
        _lib.check(lib.nrl_late_fusion_fwd(_p(hist_vec), _p(off), B, E, _p(user), _stream()), "nrl_late_fusion_fwd")
        ctx.save_for_backward(off)
        ctx.cfg = (B, E, hist_vec.shape[0])
        return user

    @staticmethod
    def backward(ctx, d_user):
        lib = _lib.load()
        (off,) = ctx.saved_tensors
        B, E, n = ctx.cfg
        d_user = d_user.contiguous().float()
        dx = torch.zeros(n, E, dtype=torch.float32, device=d_user.device)
        _lib.check(lib.nrl_late_fusion_bwd(_p(d_user), _p(off), B, E, _p(dx), _stream()), "nrl_late_fusion_bwd")
        return dx, None, None


class ScoreFn(torch.autograd.Function):
    """``DotProduct.forward`` on ragged candidates (``layers/click_predictor.py:9-11``)."""

    @staticmethod
    def forward(ctx, user, cand, off, B, Cmax):
        lib = _lib.load()
        user = _chk(user.contiguous(), torch.float32, "user")
        cand = _chk(cand.contiguous(), torch.float32, "cand")
        E = user.shape[1]
        scores = torch.empty(B, Cmax, dtype=torch.float32, device=user.device)
        _lib.check(lib.nrl_score_fwd(_p(user), _p(cand), _p(off), B, Cmax, E, _p(scores), _stream()), "nrl_score_fwd")
        ctx.save_for_backward(user, cand, off)
        ctx.cfg = (B, Cmax, E)
        return scores

    @staticmethod
    def backward(ctx, d_scores):
        lib = _lib.load()
        user, cand, off = ctx.saved_tensors
        B, Cmax, E = ctx.cfg
        d_scores = d_scores.contiguous().float()
        d_user = torch.empty_like(user)
        d_cand = torch.zeros_like(cand)
        _lib.check(lib.nrl_score_bwd(_p(d_scores), _p(user), _p(cand), _p(off), B, Cmax, E, _p(d_user),
                                     _p(d_cand), _stream()), "nrl_score_bwd")
        return d_user, d_cand, None, None, None


class CESoftFn(torch.autograd.Function):
    """``CrossEntropyLoss()(scores, y_true)`` with float targets (``nrms_module.py:277,288``)."""

    @staticmethod
    def forward(ctx, scores, labels, off):
        lib = _lib.load()
        scores = _chk(scores.contiguous(), torch.float32, "scores")
        labels = _chk(labels.contiguous(), torch.float32, "labels")
        B, Cmax = scores.shape
        loss = torch.empty(1, dtype=torch.float32, device=scores.device)
        _lib.check(lib.nrl_ce_soft_fwd(_p(scores), _p(labels), _p(off), B, Cmax, None, _p(loss), None,
                                       _stream()), "nrl_ce_soft_fwd")
        ctx.save_for_backward(scores, labels, off)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        scores, labels, off = ctx.saved_tensors
        B, Cmax = scores.shape
        g = g.contiguous().float().reshape(1)
        d = torch.empty_like(scores)
        _lib.check(lib.nrl_ce_soft_bwd(_p(scores), _p(labels), _p(off), B, Cmax, _p(g), 1.0, _p(d), _stream()),
                   "nrl_ce_soft_bwd")
        return d, None, None


class SupConFn(torch.autograd.Function):
    """``SupConLoss()`` on the score matrix (``nrms_module.py:289-316``, ``components/losses.py:6-40``) or, with
    ``dual_loss_coef`` given, the dual loss ``(1 - coef) * CE + coef * SupCon`` (``nrms_module.py:318-328``)."""

    TEMPERATURE = 0.1  # SupConLoss() default: the module's `temperature` hparam never reaches it (abstract_recommender.py:117-120)

    @staticmethod
    def forward(ctx, scores, labels, off, dual_loss_coef=None):
        lib = _lib.load()
        scores = _chk(scores.contiguous(), torch.float32, "scores")
        labels = _chk(labels.contiguous(), torch.float32, "labels")
        B, Cmax = scores.shape
        buf = torch.empty(B + 4, dtype=torch.float32, device=scores.device)  # row losses, stats[3], CE
        row, stats, ce = buf[:B], buf[B:B + 3], buf[B + 3:B + 4]
        loss = torch.empty(1, dtype=torch.float32, device=scores.device)
        if dual_loss_coef is not None:
            _lib.check(lib.nrl_ce_soft_fwd(_p(scores), _p(labels), _p(off), B, Cmax, None, _p(ce), None, _stream()),
                       "nrl_ce_soft_fwd")
        _lib.check(lib.nrl_supcon_fwd(_p(scores), _p(labels), _p(off), B, Cmax, SupConFn.TEMPERATURE,
                                      _p(ce) if dual_loss_coef is not None else None,
                                      float(dual_loss_coef or 0.0), _p(row), _p(loss), _p(stats), _stream()),
                   "nrl_supcon_fwd")
        ctx.save_for_backward(scores, labels, off, buf)
        ctx.coef = dual_loss_coef
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        scores, labels, off, buf = ctx.saved_tensors
        B, Cmax = scores.shape
        row, stats = buf[:B], buf[B:B + 3]
        g = g.contiguous().float().reshape(1)
        d = torch.empty_like(scores)
        coef = ctx.coef
        if coef is not None:
            _lib.check(lib.nrl_ce_soft_bwd(_p(scores), _p(labels), _p(off), B, Cmax, _p(g), 1.0 - float(coef), _p(d),
                                           _stream()), "nrl_ce_soft_bwd")
        _lib.check(lib.nrl_supcon_bwd(_p(scores), _p(labels), _p(off), B, Cmax, SupConFn.TEMPERATURE, _p(row), _p(stats),
                                      _p(g), 1.0 if coef is None else float(coef), 0 if coef is None else 1, _p(d),
                                      _stream()), "nrl_supcon_bwd")
        return d, None, None, None


class AdditiveFn(torch.autograd.Function):
    """``AdditiveAttention.forward`` (``layers/attention.py:24-42``): x ``[G, L, D]`` -> ``[G, D]``."""

    @staticmethod
    def forward(ctx, x, weight, bias, query, precision):
        lib = _lib.load()
        x = _chk(x.contiguous(), torch.float32, "x")
        G, L, D = x.shape
        Q = query.numel()
        out = torch.empty(G, D, dtype=torch.float32, device=x.device)
        ws = workspace(lib.nrl_additive_ws_bytes(G, L, D, Q), x.device)
        _lib.check(lib.nrl_additive_fwd(_p(x), G, L, D, Q, _p(_chk(weight, torch.float32, "weight")),
                                        _p(_chk(bias, torch.float32, "bias")), _p(_chk(query, torch.float32, "query")),
                                        _p(out), _p(ws), ws.numel(), precision, _stream()), "nrl_additive_fwd")
        ctx.save_for_backward(x, weight, bias, query)
        ctx.ws, ctx.precision = ws, precision
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        x, weight, bias, query = ctx.saved_tensors
        G, L, D = x.shape
        Q = query.numel()
        d_out = d_out.contiguous().float()
        dx = torch.empty_like(x)
        gw, gb, gq = torch.zeros_like(weight), torch.zeros_like(bias), torch.zeros_like(query)
        _lib.check(lib.nrl_additive_bwd(_p(x), G, L, D, Q, _p(weight), _p(query), _p(d_out), _p(dx), _p(gw),
                                        _p(gb), _p(gq), _p(_ws(ctx)), ctx.ws.numel(), ctx.precision, _stream()),
                   "nrl_additive_bwd")
        ctx.ws = None
        return dx, gw, gb, gq, None


def additive_attention(x: torch.Tensor, weight, bias, query, precision: int = PREC_BF16X3) -> torch.Tensor:
    return AdditiveFn.apply(x, weight, bias, query, precision)


def cnn_struct(tensors) -> CnnParams:
    s = CnnParams()
    for (name, _), t in zip(CnnParams._fields_, tensors):
        setattr(s, name, _p(_chk(t, torch.float32, name)))
    return s


class CnnEncoderFn(torch.autograd.Function):
    """``CNNAddAtt.forward`` (reference ``encoders/news/text.py:163-176``)."""

    @staticmethod
    def forward(ctx, ids, table, cnn_w, cnn_b, w_add, b_add, q_add, window, dropout_p, training, seed, precision):
        lib = _lib.load()
        _chk(ids, torch.int64, "ids"); _chk(table, torch.float32, "embedding table")
        n, L = ids.shape
        F_ = cnn_w.shape[0]
        dims = CnnDims(int(table.shape[1]), int(F_), int(window), int(q_add.numel()))
        prm = [cnn_w, cnn_b, w_add, b_add, q_add]
        need = lib.nrl_cnn_encoder_ws_bytes(n, L, dims)
        if need == 0:
            raise RuntimeError("nrl_cnn_encoder: unsupported dims (odd window, E % 4 == 0, F % 4 == 0, Q <= 256)")
        out = torch.empty(n, F_, dtype=torch.float32, device=table.device)
        ws = workspace(need, table.device)
        ps = cnn_struct(prm)
        _lib.check(lib.nrl_cnn_encoder_fwd(_p(ids), n, L, _p(table), table.shape[0], C.byref(ps), dims,
                                           float(dropout_p), int(training), int(seed), _p(out), _p(ws),
                                           ws.numel(), precision, _stream()), "nrl_cnn_encoder_fwd")
        ctx.save_for_backward(ids, table, *prm)
        ctx.ws, ctx.dims, ctx.cfg = ws, dims, (float(dropout_p), int(training), int(seed), precision)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        ids, table, *prm = ctx.saved_tensors
        n, L = ids.shape
        p, training, seed, precision = ctx.cfg
        d_out = d_out.contiguous().float()
        grads = [torch.zeros_like(t) for t in prm]
        d_table = torch.zeros_like(table)
        ps, gs = cnn_struct(prm), cnn_struct(grads)
        _lib.check(lib.nrl_cnn_encoder_bwd(_p(ids), n, L, table.shape[0], C.byref(ps), ctx.dims, p, training, seed,
                                           _p(d_out), C.byref(gs), _p(d_table), _p(_ws(ctx)), ctx.ws.numel(),
                                           precision, _stream()), "nrl_cnn_encoder_bwd")
        ctx.ws = None
        return (None, d_table, *grads, None, None, None, None, None)


class LinearEncoderFn(torch.autograd.Function):
    """``LinearEncoder.forward`` with ``linear_transform=True`` (``encoders/news/category.py:73-82``)."""

    @staticmethod
    def forward(ctx, ids, table, weight, bias, dropout_p, training, seed, precision):
        lib = _lib.load()
        ids = _chk(ids.contiguous(), torch.int64, "ids").reshape(-1)
        _chk(table, torch.float32, "category table")
        n, CE, O = ids.numel(), table.shape[1], weight.shape[0]
        out = torch.empty(n, O, dtype=torch.float32, device=table.device)
        ws = workspace(lib.nrl_linear_encoder_ws_bytes(n, CE, O), table.device)
        _lib.check(lib.nrl_linear_encoder_fwd(_p(ids), n, _p(table), table.shape[0], CE,
                                              _p(_chk(weight, torch.float32, "weight")),
                                              _p(_chk(bias, torch.float32, "bias")), O, float(dropout_p),
                                              int(training), int(seed), _p(out), _p(ws), ws.numel(), precision,
                                              _stream()), "nrl_linear_encoder_fwd")
        ctx.save_for_backward(ids, table, weight, bias, out)
        ctx.ws, ctx.cfg = ws, (float(dropout_p), int(training), int(seed), precision)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        ids, table, weight, bias, out = ctx.saved_tensors
        n, CE, O = ids.numel(), table.shape[1], weight.shape[0]
        p, training, seed, precision = ctx.cfg
        d_out = d_out.contiguous().float()
        gw, gb, d_table = torch.zeros_like(weight), torch.zeros_like(bias), torch.zeros_like(table)
        _lib.check(lib.nrl_linear_encoder_bwd(_p(ids), n, table.shape[0], CE, _p(weight), O, p, training, seed,
                                              _p(out), _p(d_out), _p(gw), _p(gb), _p(d_table), _p(_ws(ctx)),
                                              ctx.ws.numel(), precision, _stream()), "nrl_linear_encoder_bwd")
        ctx.ws = None
        return None, d_table, gw, gb, None, None, None, None


class PlmHeadFn(torch.autograd.Function):
    """The post-transformer part of ``PLM.forward`` (``encoders/news/text.py:93-100``): dropout ->
    MHSA over dim 0 of ``[N, T, E]`` (the reference's ``batch_first=False`` quirk) -> dropout ->
    additive pooling over the T tokens."""

    @staticmethod
    def forward(ctx, x, w_in, b_in, w_out, b_out, w_add, b_add, q_add, num_heads, attention_axis, dropout_p,
                training, seed, precision):
        lib = _lib.load()
        x = _chk(x.contiguous(), torch.float32, "x")
        N, T, E = x.shape
        dims = dims_of(E, num_heads, q_add.numel())
        blk = [w_in, b_in, w_out, b_out, w_add, b_add, q_add]
        need = lib.nrl_user_encoder_ws_bytes(N, T, dims)
        if need == 0:
            raise RuntimeError("nrl_plm_head: unsupported dims (E / heads must be 16, 20, 32, 48 or 64; Q <= 256)")
        out = torch.empty(N, E, dtype=torch.float32, device=x.device)
        ws = workspace(need, x.device)
        bs = block_struct(blk)
        _lib.check(lib.nrl_plm_head_fwd(_p(x), N, T, C.byref(bs), dims, int(attention_axis), float(dropout_p),
                                        int(training), int(seed), _p(out), _p(ws), ws.numel(), precision,
                                        _stream()), "nrl_plm_head_fwd")
        ctx.save_for_backward(*blk)
        ctx.ws, ctx.dims = ws, dims
        ctx.cfg = (N, T, E, int(attention_axis), float(dropout_p), int(training), int(seed), precision)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        blk = list(ctx.saved_tensors)
        N, T, E, axis, p, training, seed, precision = ctx.cfg
        d_out = d_out.contiguous().float()
        grads = [torch.zeros_like(t) for t in blk]
        d_x = torch.empty(N, T, E, dtype=torch.float32, device=d_out.device)
        bs, gs = block_struct(blk), block_struct(grads)
        _lib.check(lib.nrl_plm_head_bwd(N, T, C.byref(bs), ctx.dims, axis, p, training, seed, _p(d_out),
                                        C.byref(gs), _p(d_x), _p(_ws(ctx)), ctx.ws.numel(), precision, _stream()),
                   "nrl_plm_head_bwd")
        ctx.ws = None
        return (d_x, *grads, None, None, None, None, None, None)


# ----------------------------------------------------------------------------------------
# PLM transformer: embeddings + encoder layers of a HF RobertaModel / BertModel
# ----------------------------------------------------------------------------------------
from ._lib import TFM_EMBED_FIELDS, TFM_LAYER_FIELDS, TfmDims, TfmEmbed, TfmLayer  # noqa: E402


class TfmState:
    """Per-module state of the sm_100a transformer: the ``nrl_tfm_dims`` and the packed bf16 hi/lo GEMM operands of
    the layers' weights.  A layer is re-packed when one of its tensors was replaced or modified through torch (data
    pointer / version counter) or -- trainable layers only -- when this library stepped an optimizer through raw pointers
    since (``weights_epoch()``: ``nrl_adam_step`` on a flat buffer does not bump torch's counters).  The two encoder calls
    of a training step (history, candidates) therefore share one packing."""

    def __init__(self, hidden: int, heads: int, intermediate: int, num_layers: int, vocab: int, max_pos: int,
                 pad_idx: int, ln_eps: float, hidden_dropout: float, attn_dropout: float, position_mode: int = 0):
        """``position_mode``: 0 = RoBERTa position ids (from the non-padding tokens), 1 = BERT (0..T-1)."""
        self.dims = TfmDims(int(hidden), int(heads), int(intermediate), int(num_layers), int(vocab), int(max_pos),
                            int(pad_idx), float(ln_eps), float(hidden_dropout), float(attn_dropout), int(position_mode))
        self.wpack: Optional[torch.Tensor] = None
        self.packed_key = [None] * int(num_layers)
        self.precision = None

    def pack(self, layer_params, precision: int) -> torch.Tensor:
        lib = _lib.load()
        dev = layer_params[0][0].device
        need = lib.nrl_tfm_wpack_bytes(self.dims)
        if need == 0:
            raise RuntimeError(f"nrl_tfm: unsupported transformer dims: {_lib.load().nrl_last_error().decode()}")
        if self.wpack is None or self.wpack.device != dev or self.precision != precision:
            self.wpack = workspace(need, dev)
            self.packed_key = [None] * self.dims.num_layers
            self.precision = precision
        for l, ps in enumerate(layer_params):
            weights = [ps[i] for i, n in enumerate(TFM_LAYER_FIELDS) if not n.startswith("ln")]
            key = (weights_epoch() if any(t.requires_grad for t in weights) else -1,
                   tuple((t.data_ptr(), t._version) for t in weights))
            if key == self.packed_key[l]:
                continue
            st = TfmLayer()
            for n, t in zip(TFM_LAYER_FIELDS, ps):
                setattr(st, n, _p(_chk(t, torch.float32, n)))
            _lib.check(lib.nrl_tfm_pack_weights(C.byref(st), l, 1, self.dims, _p(self.wpack), self.wpack.numel(),
                                                precision, _stream()), "nrl_tfm_pack_weights")
            self.packed_key[l] = key
        return self.wpack


def _tfm_structs(params, n_layers):
    emb = TfmEmbed()
    for n, t in zip(TFM_EMBED_FIELDS, params[:5]):
        setattr(emb, n, _p(t))
    layers = (TfmLayer * n_layers)()
    for l in range(n_layers):
        for i, n in enumerate(TFM_LAYER_FIELDS):
            setattr(layers[l], n, _p(params[5 + 16 * l + i]))
    return emb, layers


class TfmEncoderFn(torch.autograd.Function):
    """``self.plm_model(**text)[0]`` of ``PLM.forward`` (``encoders/news/text.py:92``) for a RoBERTa / BERT-shaped
    transformer: ``input_ids``, ``attention_mask`` int64 ``[N, T]`` -> last hidden state ``[N, T, D]``.
    ``params``: the 5 embedding tensors (``TFM_EMBED_FIELDS`` order; ``type0`` = row 0 of the token-type table) then 16
    tensors per layer (``TFM_LAYER_FIELDS`` order).  A layer with no trainable tensor is frozen (data gradient only);
    mixed layers are refused by the library."""

    @staticmethod
    def forward(ctx, input_ids, attention_mask, state, training, seed, precision, *params):
        lib = _lib.load()
        ids = _chk(input_ids.contiguous(), torch.int64, "input_ids")
        mask = None if attention_mask is None else _chk(attention_mask.contiguous(), torch.int64, "attention_mask")
        N, T = ids.shape
        dims = state.dims
        L = dims.num_layers
        if len(params) != 5 + 16 * L:
            raise RuntimeError(f"TfmEncoderFn: expected {5 + 16 * L} parameter tensors, got {len(params)}")
        params = [_chk(t.contiguous(), torch.float32, "transformer parameter") for t in params]
        wpack = state.pack([params[5 + 16 * l: 21 + 16 * l] for l in range(L)], precision)
        # no gradient will be asked for (torch.no_grad(), or nothing trainable): the layers share one set of activations
        keep = int(any(ctx.needs_input_grad[6:]))
        need = lib.nrl_tfm_ws_bytes(N, T, dims, keep)
        if need == 0:
            raise RuntimeError("nrl_tfm: unsupported transformer dims")
        ws = workspace(need, ids.device)
        out = torch.empty(N, T, dims.hidden, dtype=torch.float32, device=ids.device)
        emb, layers = _tfm_structs(params, L)
        _lib.check(lib.nrl_tfm_encoder_fwd(_p(ids), _p(mask), N, T, C.byref(emb), layers, dims, int(training), int(seed),
                                           _p(wpack), _p(out), keep, _p(ws), ws.numel(), precision, _stream()),
                   "nrl_tfm_encoder_fwd")
        if not keep:
            return out
        ctx.save_for_backward(ids, *params)
        ctx.mask, ctx.ws, ctx.wpack, ctx.state = mask, ws, wpack, state
        ctx.cfg = (N, T, int(training), int(seed), precision)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        ids, *params = ctx.saved_tensors
        N, T, training, seed, precision = ctx.cfg
        dims = ctx.state.dims
        L = dims.num_layers
        need = ctx.needs_input_grad[6:]
        d_out = d_out.contiguous().float()
        grads = [torch.zeros_like(t) if nd else None for t, nd in zip(params, need)]
        emb, layers = _tfm_structs(params, L)
        eg = None
        if any(need[:5]):
            if not all(need[:5]):
                raise RuntimeError("TfmEncoderFn: the embeddings must be all trainable or all frozen")
            eg = TfmEmbed()
            for n, t in zip(TFM_EMBED_FIELDS, grads[:5]):
                setattr(eg, n, _p(t))
        lg = (TfmLayer * L)()
        for l in range(L):
            for i, n in enumerate(TFM_LAYER_FIELDS):
                setattr(lg[l], n, _p(grads[5 + 16 * l + i]))
        _lib.check(lib.nrl_tfm_encoder_bwd(_p(ids), _p(ctx.mask), N, T, C.byref(emb), layers, dims, training, seed,
                                           _p(ctx.wpack), _p(d_out), C.byref(eg) if eg is not None else None, lg,
                                           _p(_ws(ctx)), ctx.ws.numel(), precision, _stream()), "nrl_tfm_encoder_bwd")
        ctx.ws = None
        return (None, None, None, None, None, None, *grads)


def tfm_hidden_dropout_mask(R: int, D: int, site: int, seed: int, p: float, device) -> torch.Tensor:
    """keep flags ``[R, D]`` (bool) of hidden-dropout site ``site`` (0 embeddings, 1 + 2l attention output, 2 + 2l layer
    output of layer l) that ``TfmEncoderFn`` uses for ``seed``: lets a test replay the kernel's masks in the oracle."""
    keep = torch.empty(R, D, dtype=torch.uint8, device=device)
    _lib.check(_lib.load().nrl_tfm_hidden_dropout_mask(_p(keep), R, D, int(site), int(seed), float(p), _stream()),
               "nrl_tfm_hidden_dropout_mask")
    return keep.bool()


def tfm_attn_dropout_mask(layers: int, N: int, heads: int, T: int, seed: int, p: float, device) -> torch.Tensor:
    """keep flags ``[layers, N, heads, T, T]`` (bool) of the attention-probability dropout."""
    lib = _lib.load()
    keep = torch.empty(layers, N, heads, T, T, dtype=torch.uint8, device=device)
    for l in range(layers):
        for n in range(N):
            for h in range(heads):
                _lib.check(lib.nrl_tfm_attn_dropout_mask(_p(keep[l, n, h]), l, n, h, heads, T, int(seed), float(p),
                                                         _stream()), "nrl_tfm_attn_dropout_mask")
    return keep.bool()
