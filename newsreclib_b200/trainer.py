"""Fused NRMS training step on one GPU (one process per GPU under ``torch.distributed``).

``NRMSTrainer`` owns ONE flat fp32 parameter buffer (embedding table, title block, user block,
in the reference's ``state_dict`` order and names), one flat gradient buffer and
the Adam moments.  A step is three library calls on the current CUDA stream:

    nrl_nrms_step (forward + CE + backward, ~45 kernels)
    [-> one NCCL all-reduce of the flat gradient buffer when world_size > 1]
    ->  nrl_adam_step_zero_grad (dense Adam over the flat buffer, torch.optim.Adam semantics, which also
        clears the gradient buffer for the next step: optimizer.zero_grad() costs no extra pass)

or, with ``exchange="peer"``, the last two lines are ONE kernel over NVLink peer memory
(``nrl_exchange_adam_step``: reduce-scatter by peer loads, Adam on the owned slice, all-gather by
peer stores; ``exchange.py``).

The path shards by impression batch (each rank runs its own ``batch_size`` impressions, as
Lightning DDP does for the reference, ``configs/trainer/ddp.yaml``); the only exchange is the
gradient all-reduce, averaged by ``grad_scale = 1 / world_size`` inside the Adam kernel.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

from . import _lib, ops

TITLE = "news_encoder.text_encoders.title."
USER = "user_encoder."


class FlatParams:
    """ONE flat fp32 buffer holding every parameter under the reference's ``state_dict`` names,
    plus same-shaped gradient / Adam-moment buffers.  Pure host-side bookkeeping (works on any
    device, so the world_size-2 gloo tests exercise it on CPU).  Every tensor starts on a
    16-byte boundary (the table gradient uses 128-bit atomics)."""

    def __init__(self, params: Dict[str, torch.Tensor], keys, device, buffers=None) -> None:
        """``buffers(total) -> (flat, grad)`` lets the caller place the parameter and gradient buffers
        (zeroed, ``total`` fp32 each) in memory of its choice -- the peer-mapped block of
        ``exchange.PeerBlock`` for the fused NVLink exchange; default: ordinary torch tensors."""
        self.keys = list(keys)
        shapes = [tuple(params[k].shape) for k in self.keys]
        sizes = [params[k].numel() for k in self.keys]
        self.offsets, total = [], 0
        for n in sizes:
            self.offsets.append(total)
            total += (n + 3) // 4 * 4
        if buffers is None:
            self.flat = torch.zeros(total, dtype=torch.float32, device=device)
            self.grad = torch.zeros_like(self.flat)
        else:
            self.flat, self.grad = buffers(total)
            if self.flat.numel() != total or self.grad.numel() != total:
                raise ValueError("FlatParams: buffers() must return two tensors of `total` elements")
        self.m = torch.zeros(total, dtype=torch.float32, device=device)
        self.v = torch.zeros_like(self.m)
        self.params, self.grads = {}, {}
        for k, o, n, shp in zip(self.keys, self.offsets, sizes, shapes):
            self.params[k] = self.flat[o:o + n].view(shp)
            self.grads[k] = self.grad[o:o + n].view(shp)
            self.params[k].copy_(params[k])

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: v.detach().clone() for k, v in self.params.items()}


class GradExchange:
    """The path's single exchange step: sum the flat gradient buffer over the ranks (NCCL over
    NVLink on GPUs, gloo in the CPU tests) and hand back the ``1 / world_size`` factor the Adam
    kernel folds into its gradient read — together the mean Lightning DDP applies for the
    reference (``configs/trainer/ddp.yaml``).  world_size 1 = no-op."""

    def __init__(self, process_group=None) -> None:
        self.pg = process_group
        on = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.world = torch.distributed.get_world_size(process_group) if on else 1

    def all_reduce(self, flat_grad: torch.Tensor) -> float:
        if self.world > 1:
            torch.distributed.all_reduce(flat_grad, op=torch.distributed.ReduceOp.SUM, group=self.pg)
        return 1.0 / self.world

    def all_reduce_chunks(self, flat_grad: torch.Tensor, chunk_elems: int = 0):
        """Same exchange, pipelined: the flat buffer is all-reduced in chunks on the communicator's stream
        (async), and ``(slice, work)`` pairs are yielded in order so the caller can run the optimizer on
        chunk i while chunks i+1.. are still on the wire.  world_size 1 yields the whole buffer at once."""
        n = flat_grad.numel()
        if self.world == 1:
            yield slice(0, n), None
            return
        if chunk_elems <= 0:  # default: ONE all-reduce of the whole buffer (measured on 2 x B200: 2.47 ms/step
            # against 2.52 / 2.69 with 16 MB / 8 MB chunks -- per-collective latency beats the Adam overlap);
            # NRL_EXCHANGE_CHUNK_MB > 0 pipelines in chunks of that size
            mb = float(os.environ.get("NRL_EXCHANGE_CHUNK_MB", "0"))
            chunk_elems = n if mb <= 0 else int(mb * (1 << 20) / 4)
        chunk_elems = max(4, chunk_elems // 4 * 4)  # 16-byte aligned slices (the Adam kernel's vector path)
        works = []
        for lo in range(0, n, chunk_elems):
            sl = slice(lo, min(n, lo + chunk_elems))
            works.append((sl, torch.distributed.all_reduce(flat_grad[sl], op=torch.distributed.ReduceOp.SUM,
                                                           group=self.pg, async_op=True)))
        for sl, work in works:
            work.wait()  # the current stream waits for this chunk only
            yield sl, work

    def gather_sharded(self, t: torch.Tensor) -> torch.Tensor:
        """Full copy of a tensor that every rank holds only on the slice it owns and as zeros elsewhere (the Adam
        moments under the fused peer exchange, ``exchange.slice_bounds``): the sum over the ranks."""
        out = t.detach().clone()
        if self.world > 1:
            torch.distributed.all_reduce(out, op=torch.distributed.ReduceOp.SUM, group=self.pg)
        return out

    @staticmethod
    def rank_seed(base_seed: int, rank: int, step: int) -> int:
        """Dropout seed of (rank, step): ranks must not share masks (they see different
        impressions), and a rank must never reuse a seed across steps."""
        return (int(base_seed) + 1000003 * int(rank) + int(step)) & ((1 << 62) - 1)


class NRMSTrainer:
    def __init__(self, params: Dict[str, torch.Tensor], num_heads: int, *, device="cuda",
                 dropout_p: float = 0.2, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 precision: int = ops.PREC_BF16X3, late_fusion: bool = False, seed: int = 1234,
                 process_group=None, exchange: Optional[str] = None, exchange_timeout_s: float = 30.0,
                 status_every: int = 64) -> None:
        """``exchange``: how the ranks' gradients meet the optimizer when world_size > 1 --
        ``"nccl"`` (one NCCL all-reduce of the flat gradient buffer, then dense Adam on every rank) or
        ``"peer"`` (``nrl_exchange_adam_step``: reduce-scatter + sharded Adam + all-gather as one kernel
        over NVLink peer memory, ``exchange.py``).  Default: ``$NRL_EXCHANGE`` or ``"nccl"``.
        ``exchange_timeout_s``: how long a rank's fused exchange kernel waits for a peer before it gives up (a dead
        peer must not hang the GPU); raise it if one rank may legitimately stall for longer between two steps
        (checkpointing on rank 0, a slow data loader).
        ``status_every``: every that many steps (and in ``state_dict`` / ``gather_moments``) the stream is synchronised
        and the library's sticky error words are read -- a timed-out peer barrier or an out-of-range token / segment
        id raises ``RuntimeError`` here instead of silently training on (0 = only at checkpoints)."""
        _lib.load()
        self.device = torch.device(device)
        self.keys = [TITLE + "embedding_layer.weight"] + [TITLE + k for k in ops.BLOCK_KEYS] + \
                    [USER + k for k in ops.BLOCK_KEYS]
        self.exchange_mode = (exchange or os.environ.get("NRL_EXCHANGE", "nccl")).lower()
        if self.exchange_mode not in ("nccl", "peer"):
            raise ValueError(f"exchange must be 'nccl' or 'peer', got {self.exchange_mode!r}")
        self.peer_block = None
        buffers = None
        if self.exchange_mode == "peer":
            from .exchange import PeerBlock

            def buffers(total):
                # the embedding table leads the flat buffer: its gradient is row-sparse (a step touches the tokens
                # of its batch), so the exchange moves only the rows a rank actually touched
                table = params[self.keys[0]]
                rows, width = (table.shape[0], table.shape[1]) if table.shape[1] % 4 == 0 else (0, 0)
                self.peer_block = PeerBlock(total, self.device, process_group, sparse_rows=rows, row_elems=width)
                return self.peer_block.flat, self.peer_block.grad
        self.fp = FlatParams(params, self.keys, self.device, buffers)
        self.flat, self.grad, self.m, self.v = self.fp.flat, self.fp.grad, self.fp.m, self.fp.v
        self.params, self.grads = self.fp.params, self.fp.grads
        self.table = self.params[self.keys[0]]
        E = self.table.shape[1]
        Q = self.params[TITLE + "additive_attention.query"].numel()
        self.dims = ops.dims_of(E, num_heads, Q)
        self.news_block = ops.block_from_dict(self.params, TITLE)
        self.user_block = ops.block_from_dict(self.params, USER)
        self.grad_pack = (ops.block_from_dict(self.grads, TITLE), ops.block_from_dict(self.grads, USER),
                          self.grads[self.keys[0]])
        self.dropout_p, self.lr, self.betas, self.eps = dropout_p, lr, betas, eps
        self.precision, self.late_fusion, self.seed = precision, late_fusion, seed
        self.step_count = 0
        self.exchange_epoch = 0
        self.exchange_ctas = int(os.environ.get("NRL_EXCHANGE_CTAS", "0"))  # grid of the fused exchange; 0 = 4 per SM
        self.exchange_timeout_s = float(exchange_timeout_s)
        self.ws: Optional[torch.Tensor] = None
        self.exchange = GradExchange(process_group)
        self.world = self.exchange.world
        on = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.rank = torch.distributed.get_rank(process_group) if on else 0
        self._structs = None
        self.status_every = int(status_every)
        self._grads_clean = True  # the gradient buffers are zero (fresh, or cleared by the fused Adam step)
        if self.world > 1:
            # what Lightning DDP does for the reference at construction: every replica starts from rank 0's values
            # (ranks built from different seeds / checkpoints would otherwise average gradients of different weights)
            torch.distributed.broadcast(self.flat, src=torch.distributed.get_global_rank(process_group, 0)
                                        if process_group is not None else 0, group=process_group)

    def check_status(self) -> None:
        """Synchronise the stream and raise if a peer-exchange barrier timed out (the replica of that step may be
        partially updated: restore from the last checkpoint) or a device-side input check fired."""
        if self.peer_block is not None:
            code = self.peer_block.status()
            if code:
                raise RuntimeError(f"rank {self.rank}: peer-exchange barrier timed out (code {code}: "
                                   f"{'ready' if code == 1 else 'done'} wait; a peer never arrived within "
                                   f"{self.exchange_timeout_s:.0f} s). Parameters after step {self.step_count} are not "
                                   "trustworthy; restore the last checkpoint on all ranks.")
        if self.device.type == "cuda":
            with torch.cuda.device(self.device):
                ops.device_status(raise_on_error=True)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Reference-named parameters (loadable into the reference's NRMSModule)."""
        self.check_status()
        return {k: v.detach().clone() for k, v in self.params.items()}

    def gather_moments(self):
        """Full Adam moments ``(exp_avg, exp_avg_sq)`` as flat tensors in ``FlatParams`` order, on every rank (for an
        optimizer checkpoint).  With the fused peer exchange a rank updates only the slice it owns
        (``exchange.slice_bounds``) and the rest stays zero, so the full state is the sum over the ranks; with the
        NCCL exchange every rank already holds it."""
        self.check_status()
        if self.peer_block is None:
            return self.m.detach().clone(), self.v.detach().clone()
        return self.exchange.gather_sharded(self.m), self.exchange.gather_sharded(self.v)

    # ------------------------------------------------------------------ steps
    def _finish(self) -> None:
        scale = 1.0 / self.world
        self.step_count += 1
        if self.peer_block is not None:
            from .exchange import exchange_adam_step
            self.exchange_epoch += 1  # barrier epoch: never reset, also when step_count is
            exchange_adam_step(self.peer_block.peer_set, self.m, self.v, self.flat.numel(), self.step_count,
                               lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, grad_scale=scale,
                               epoch=self.exchange_epoch, max_ctas=self.exchange_ctas, timeout_s=self.exchange_timeout_s,
                               sparse_rows=self.peer_block.sparse_rows, row_elems=self.peer_block.row_elems,
                               zero_grads=True)
            self._grads_clean = True  # every owner cleared what it consumed, on every rank, before the done barrier
        else:
            # Adam on chunk i overlaps the all-reduce of chunks i+1.. (one chunk = everything when world == 1);
            # the same pass clears the gradient slice it has consumed
            for sl, _ in self.exchange.all_reduce_chunks(self.grad):
                ops.adam_step(self.flat[sl], self.grad[sl], self.m[sl], self.v[sl], self.step_count, self.lr,
                              self.betas[0], self.betas[1], self.eps, grad_scale=scale, zero_grad=True)
            self._grads_clean = True
        if self.status_every > 0 and self.step_count % self.status_every == 0:
            self.check_status()

    def _zero_grads(self) -> None:
        if not self._grads_clean:
            self.grad.zero_()
        self._grads_clean = False

    def train_step(self, batch: Dict, B: int, Hmax: int, Cmax: int, training: bool = True):
        """Device-resident batch -> (scores [B, Cmax], loss [1]) device tensors; no host sync."""
        self._zero_grads()
        scores, loss, self.ws = ops.nrms_step(
            batch, self.table, self.news_block, self.user_block, self.dims, B=B, Hmax=Hmax, Cmax=Cmax,
            late_fusion=self.late_fusion, dropout_p=self.dropout_p, training=training,
            seed=GradExchange.rank_seed(self.seed, self.rank, self.step_count), grads=self.grad_pack, ws=self.ws, precision=self.precision)
        self._finish()
        return scores, loss

    def probe_exchange(self, batch: Dict, B: int, Hmax: int, Cmax: int):
        """One training step that ALSO checks the exchange against the library path it replaces: forward + backward into
        the (clean) gradient buffer, a copy of this rank's gradients is summed with NCCL and fed to a dense Adam step on a
        copy of the parameters, then the configured exchange runs.  Returns ``(equals, cleared)``: the updated replica
        equals Adam on the averaged gradients (a parameter whose gradient is pure rounding noise -- the key third of
        ``in_proj_bias`` is mathematically zero -- moves by +-lr in either summation order; everything else must agree to
        fp32 rounding), and the gradient buffer was left cleared.  Collective: call on every rank."""
        keep = self.flat.clone()
        m0, v0, step0 = self.m.clone(), self.v.clone(), self.step_count
        self._zero_grads()
        _, _, self.ws = ops.nrms_step(batch, self.table, self.news_block, self.user_block, self.dims, B=B, Hmax=Hmax, Cmax=Cmax,
                                      late_fusion=self.late_fusion, dropout_p=self.dropout_p, training=True,
                                      seed=GradExchange.rank_seed(self.seed, self.rank, self.step_count),
                                      grads=self.grad_pack, ws=self.ws, precision=self.precision)
        g_sum = self.grad.clone()
        self._finish()
        if self.world > 1:
            torch.distributed.all_reduce(g_sum, op=torch.distributed.ReduceOp.SUM, group=self.exchange.pg)
        if self.peer_block is not None:  # moments are sharded: the reference copy needs the full ones of the start state
            m0, v0 = self.exchange.gather_sharded(m0), self.exchange.gather_sharded(v0)
        ops.adam_step(keep, g_sum, m0, v0, step0 + 1, self.lr, self.betas[0], self.betas[1], self.eps,
                      grad_scale=1.0 / self.world)
        d = (keep - self.flat).abs()
        equals = float((d > 1e-6).float().mean()) < 2e-3 and float(d.median()) < 1e-7
        return equals, not bool(self.grad.any())

    def train_step_host(self, hb: Dict, B: int, Hmax: int, Cmax: int, scores_host: torch.Tensor,
                        loss_host: torch.Tensor, training: bool = True) -> None:
        """End-to-end step from HOST buffers: ``nrl_nrms_step_host_begin`` copies ids / segment ids / labels host->device
        on a copy stream and enqueues the step and the copies of scores and loss back; the gradient exchange and the
        Adam update are enqueued behind it; only then does ``nrl_nrms_step_host_end`` wait for THIS step's scores and
        loss.  The optimizer step therefore runs while the host waits, returns and stages the next batch (whose copies
        overlap it on the copy stream).  ``NRL_E2E_SPLIT=0`` selects the one-call ``nrl_nrms_step_host`` (copy, step,
        copy back, synchronise, then the optimizer step)."""
        lib = _lib.load()
        hist_ids, cand_ids = hb["x_hist"]["title"], hb["x_cand"]["title"]
        nh, L = hist_ids.shape
        nc = cand_ids.shape[0]
        need = lib.nrl_nrms_ws_bytes(nh, nc, L, B, Hmax, Cmax, self.dims)
        if self.ws is None or self.ws.numel() < need:
            self.ws = ops.workspace(need, self.device)
        self._zero_grads()
        nb, ub = ops.block_struct(self.news_block), ops.block_struct(self.user_block)
        ng, ug = ops.block_struct(self.grad_pack[0]), ops.block_struct(self.grad_pack[1])
        args = (hist_ids.data_ptr(), cand_ids.data_ptr(), hb["batch_hist"].data_ptr(), hb["batch_cand"].data_ptr(),
                hb["labels"].data_ptr(), nh, nc, L, B, Hmax, Cmax, self.table.data_ptr(), self.table.shape[0],
                C.byref(nb), C.byref(ub), self.dims, int(self.late_fusion), float(self.dropout_p), int(training),
                GradExchange.rank_seed(self.seed, self.rank, self.step_count), scores_host.data_ptr(),
                loss_host.data_ptr(), 1, C.byref(ng), C.byref(ug), self.grad_pack[2].data_ptr(), self.ws.data_ptr(),
                self.ws.numel(), self.precision)
        stream = torch.cuda.current_stream().cuda_stream
        if os.environ.get("NRL_E2E_SPLIT", "1") == "0":
            _lib.check(lib.nrl_nrms_step_host(*args, stream), "nrl_nrms_step_host")
            self._finish()
            return
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        ticket = C.c_void_p()
        _lib.check(lib.nrl_nrms_step_host_begin(*args, self._copy_stream.cuda_stream, self._status_host.data_ptr(),
                                                stream, C.byref(ticket)), "nrl_nrms_step_host_begin")
        try:
            self._finish()
        finally:
            _lib.check(lib.nrl_nrms_step_host_end(ticket), "nrl_nrms_step_host_end")

    @torch.no_grad()
    def eval_forward(self, batch: Dict, B: int, Hmax: int, Cmax: int):
        scores, loss, self.ws = ops.nrms_step(
            batch, self.table, self.news_block, self.user_block, self.dims, B=B, Hmax=Hmax, Cmax=Cmax,
            late_fusion=self.late_fusion, dropout_p=0.0, training=False, ws=self.ws, precision=self.precision)
        return scores, loss


class ModuleTrainer:
    """Training loop body for any of the drop-in LightningModule mirrors (``NRMSModule``, ``NAMLModule``)
    without Lightning: ``model_step`` (sm_100a forward + loss), autograd backward through the
    ``torch.autograd.Function``s of ``ops.py``, then ONE gradient all-reduce and ONE ``nrl_adam_step`` over
    flat buffers the module's parameters and ``.grad``s are re-pointed into (what Lightning's DDP +
    ``torch.optim.Adam`` do for the reference, ``configs/model/naml.yaml:56-59``)."""

    def __init__(self, module: torch.nn.Module, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group=None, exchange: Optional[str] = None, exchange_timeout_s: float = 30.0) -> None:
        """``exchange``: ``"nccl"`` (one all-reduce of the flat gradient buffer + dense Adam on every rank) or ``"peer"``
        (``nrl_exchange_adam_step`` over NVLink peer memory, as in ``NRMSTrainer``); default ``$NRL_EXCHANGE`` or
        ``"nccl"``.  The largest ``nn.Embedding`` table leads the flat buffer so that the peer exchange can treat its
        gradient as row-sparse."""
        _lib.load()
        self.module = module
        uniq, seen = [], set()
        for p in module.parameters():
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p)); uniq.append(p)
        if not uniq or not uniq[0].is_cuda:
            raise RuntimeError("ModuleTrainer needs a module on a CUDA device (newsreclib_b200 has no CPU path)")
        tables = [m.weight for m in module.modules() if isinstance(m, torch.nn.Embedding) and m.weight.requires_grad
                  and m.weight.dim() == 2 and m.weight.shape[1] % 4 == 0]
        table = max(tables, key=lambda t: t.numel()) if tables else None
        if table is not None:
            uniq = [table] + [p for p in uniq if p is not table]
        total, offs = 0, []
        for p in uniq:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        dev = uniq[0].device
        self.exchange = GradExchange(process_group)
        self.world = self.exchange.world
        self.exchange_mode = (exchange or os.environ.get("NRL_EXCHANGE", "nccl")).lower()
        if self.exchange_mode not in ("nccl", "peer"):
            raise ValueError(f"exchange must be 'nccl' or 'peer', got {self.exchange_mode!r}")
        self.peer_block = None
        if self.exchange_mode == "peer":
            from .exchange import PeerBlock
            rows, width = (table.shape[0], table.shape[1]) if table is not None else (0, 0)
            self.peer_block = PeerBlock(total, dev, process_group, sparse_rows=rows, row_elems=width)
            self.flat, self.grad = self.peer_block.flat, self.peer_block.grad
        else:
            self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
            self.grad = torch.zeros_like(self.flat)
        self.m, self.v = torch.zeros(total, dtype=torch.float32, device=dev), torch.zeros(total, dtype=torch.float32, device=dev)
        for p, o in zip(uniq, offs):
            n = p.numel()
            self.flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + n].view_as(p)
            p.grad = self.grad[o:o + n].view_as(p)  # autograd accumulates in place into the flat buffer
        if hasattr(module, "grad_targets"):
            # modules whose model_step is one fused autograd node (NRMSModule -> ops.NrmsStepFn) let the kernels
            # accumulate straight into these views: no zeroed copy of the table gradient, no AccumulateGrad pass
            module.grad_targets = "param.grad"
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        self.exchange_epoch = 0
        self.exchange_timeout_s = float(exchange_timeout_s)
        if self.world > 1:  # Lightning DDP broadcasts rank 0's parameters at construction
            torch.distributed.broadcast(self.flat, src=torch.distributed.get_global_rank(process_group, 0)
                                        if process_group is not None else 0, group=process_group)

    def check_status(self) -> None:
        if self.peer_block is not None and self.peer_block.status():
            raise RuntimeError(f"peer-exchange barrier timed out (code {self.peer_block.status()}); restore the last checkpoint")
        ops.device_status(raise_on_error=True)

    def train_step(self, batch) -> torch.Tensor:
        self.module.train()
        loss = self.module.model_step(batch)[0]  # gradient buffers are clean: zeroed at construction / by the Adam pass
        loss.backward()
        self.step_count += 1
        if self.peer_block is not None:
            from .exchange import exchange_adam_step
            self.exchange_epoch += 1
            exchange_adam_step(self.peer_block.peer_set, self.m, self.v, self.flat.numel(), self.step_count, lr=self.lr,
                               beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, grad_scale=1.0 / self.world,
                               epoch=self.exchange_epoch, timeout_s=self.exchange_timeout_s,
                               sparse_rows=self.peer_block.sparse_rows, row_elems=self.peer_block.row_elems, zero_grads=True)
        else:
            scale = self.exchange.all_reduce(self.grad)
            ops.adam_step(self.flat, self.grad, self.m, self.v, self.step_count, self.lr, self.betas[0], self.betas[1],
                          self.eps, grad_scale=scale, zero_grad=True)
        return loss.detach()
