"""Peer-memory gradient exchange fused with the Adam step (``nrl_exchange_adam_step``).

What Lightning DDP + ``torch.optim.Adam`` do for the reference after every backward pass
(``configs/trainer/ddp.yaml``, ``configs/model/nrms.yaml:49-52``: gradient mean over the ranks, then
Adam on every replica) is ONE kernel per rank here: the flat parameter and gradient buffers of a
rank live in a ``cudaMalloc`` block that every peer maps over NVLink (CUDA IPC); a rank sums the
gradients of its 1/world slice by peer loads, runs Adam on that slice, stores the new values
into every replica and clears the gradients it has consumed; flag barriers in the same peer memory bracket the
kernel.  The embedding-table part of the gradient is row-sparse (a step touches a fraction of the vocabulary):
every rank publishes one bit per row and all-zero rows never cross the links.  ``torch.distributed``
is used once, at construction, to pass the 64-byte IPC handles around.

``slice_bounds`` is the ownership rule of the kernel restated for the host (tests, checkpoint code
that wants to gather the sharded Adam moments).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch

from . import _lib


def slice_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Element range ``[lo, hi)`` of the flat buffers that ``rank`` owns (16-byte granules, the last
    ranks may own less or nothing)."""
    if n % 4:
        raise ValueError(f"flat buffer length {n} is not a multiple of 4")
    n4 = n // 4
    per = (n4 + world - 1) // world
    lo = min(n4, rank * per)
    hi = min(n4, lo + per)
    return 4 * lo, 4 * hi


class _DevMem:
    """``__cuda_array_interface__`` view of raw device memory so that torch can alias it."""

    def __init__(self, ptr: int, nelem: int, typestr: str) -> None:
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": typestr, "data": (ptr, False),
                                         "version": 3, "strides": None}


class PeerBlock:
    """One rank's peer-mapped block: ``[ params n f32 | grads n f32 | flag block | row bitmaps ]`` and the same
    blocks of all peers opened through their IPC handles.  ``sparse_rows`` > 0 declares the first
    ``sparse_rows * row_elems`` elements a row-sparse gradient (the embedding table) and reserves
    ``world x ceil(sparse_rows / 32)`` u32 words of bitmap area."""

    # two seams for the world_size-2 gloo tests of the construction protocol (tests/test_dist_cpu.py)
    def _device_ctx(self):
        return torch.cuda.device(self.device)

    def _alias(self, ptr: int, nelem: int) -> torch.Tensor:
        return torch.as_tensor(_DevMem(ptr, nelem, "<f4"), device=self.device)

    def __init__(self, n: int, device: torch.device, process_group=None, lib=None, sparse_rows: int = 0,
                 row_elems: int = 0) -> None:
        lib = lib or _lib.load()
        dist = torch.distributed
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(process_group) if on else 1
        self.rank = dist.get_rank(process_group) if on else 0
        if self.world > _lib.MAX_RANKS:
            raise ValueError(f"peer exchange is built for <= {_lib.MAX_RANKS} ranks, got {self.world}")
        if n % 4:
            raise ValueError("flat buffer length must be a multiple of 4 (16-byte slices)")
        self.n, self.device = n, torch.device(device)
        self._lib = lib
        self._flag_off = 8 * n
        if sparse_rows and (row_elems <= 0 or row_elems % 4 or sparse_rows * row_elems > n):
            raise ValueError("sparse region must be sparse_rows rows of row_elems (multiple of 4) elements inside n")
        self.sparse_rows, self.row_elems = int(sparse_rows), int(row_elems)
        self.bm_words = (self.sparse_rows + 31) // 32
        self._bm_off = self._flag_off + _lib.FLAG_BYTES
        nbytes = self._bm_off + (4 * self.world * self.bm_words + 15) // 16 * 16
        base, handle = C.c_void_p(), C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        self.base, self._opened = 0, []
        # A rank that fails (no peer access, IPC not permitted, out of memory) must not leave the others waiting in a
        # collective: every rank goes through the same two object all-gathers, then all raise or none does.
        err: Optional[str] = None
        try:
            with self._device_ctx():
                _lib.check(lib.nrl_peer_alloc(nbytes, C.byref(base), handle), "nrl_peer_alloc")
            self.base = int(base.value)
        except Exception as e:  # noqa: BLE001
            err = f"rank {self.rank}: {e}"
        bases = [0] * self.world
        bases[self.rank] = self.base
        if self.world > 1:
            handles: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(handles, handle.raw if err is None else None, group=process_group)
            if err is None and all(h is not None for h in handles):
                try:
                    with self._device_ctx():
                        for r, h in enumerate(handles):
                            if r == self.rank:
                                continue
                            ptr = C.c_void_p()
                            _lib.check(lib.nrl_peer_open(h, C.byref(ptr)), f"nrl_peer_open(rank {r})")
                            bases[r] = int(ptr.value)
                            self._opened.append(bases[r])
                except Exception as e:  # noqa: BLE001
                    err = f"rank {self.rank}: {e}"
            errs: List[Optional[str]] = [None] * self.world
            dist.all_gather_object(errs, err, group=process_group)
            failed = [e for e in errs if e is not None] or \
                     ([f"rank {r}: allocation failed" for r, h in enumerate(handles) if h is None])
            if failed:
                self.close()
                raise RuntimeError("peer exchange unavailable: " + "; ".join(failed))
            dist.barrier(group=process_group)  # every block is zeroed and mapped before anyone signals
        elif err is not None:
            raise RuntimeError("peer exchange unavailable: " + err)
        self.flat = self._alias(self.base, n)
        self.grad = self._alias(self.base + 4 * n, n)
        self.peer_set = peer_set(self.world, self.rank, [b for b in bases], [b + 4 * n for b in bases],
                                 [b + self._flag_off for b in bases],
                                 [b + self._bm_off for b in bases] if self.sparse_rows else None)

    @property
    def flags_ptr(self) -> int:
        return self.base + self._flag_off

    def status(self) -> int:
        """Synchronise the current stream and return the flag block's error word (0 = ok)."""
        err = C.c_ulonglong(0)
        _lib.check(self._lib.nrl_exchange_status(self.flags_ptr, C.byref(err),
                                                 torch.cuda.current_stream(self.device).cuda_stream),
                   "nrl_exchange_status")
        return int(err.value)

    def timeline_us(self):
        """Phases of the LAST exchange kernel of this rank in microseconds (device ``%globaltimer``): row scan, wait for
        the slowest rank (ready barrier), owned slice (pull + Adam + push + clear), done barrier.  Synchronises."""
        torch.cuda.synchronize(self.device)
        words = torch.as_tensor(_DevMem(self.flags_ptr, _lib.FLAG_BYTES // 8, "<i8"), device=self.device)[40:45].tolist()
        t0, t1, t2, t3, t4 = words
        return {"scan": (t1 - t0) / 1e3, "ready_wait": (t2 - t1) / 1e3, "slice": (t3 - t2) / 1e3, "done_wait": (t4 - t3) / 1e3,
                "total": (t4 - t0) / 1e3}

    def close(self) -> None:
        """Unmap the peers' blocks and free the local one (call on every rank, after a barrier)."""
        with self._device_ctx():
            if (self.base or self._opened) and self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            for ptr in self._opened:
                self._lib.nrl_peer_close(ptr)
            self._opened = []
            self.flat = self.grad = None
            if self.base:
                self._lib.nrl_peer_free(self.base)
        self.base = 0


def peer_set(world: int, rank: int, params: List[int], grads: List[int], flags: List[int],
             bitmaps: Optional[List[int]] = None) -> "_lib.PeerSet":
    ps = _lib.PeerSet()
    ps.world, ps.rank = int(world), int(rank)
    for r in range(world):
        ps.params[r], ps.grads[r], ps.flags[r] = params[r], grads[r], flags[r]
        if bitmaps is not None:
            ps.bitmaps[r] = bitmaps[r]
    return ps


def exchange_adam_step(ps: "_lib.PeerSet", m: torch.Tensor, v: torch.Tensor, n: int, step: int, *, lr=1e-4,
                       beta1=0.9, beta2=0.999, eps=1e-8, grad_scale: Optional[float] = None, epoch: Optional[int] = None,
                       max_ctas: int = 0, timeout_s: float = 5.0, stream: Optional[int] = None, sparse_rows: int = 0,
                       row_elems: int = 0, zero_grads: bool = False) -> None:
    """Enqueue one fused exchange + Adam step (see ``include/nrl.h``).  ``sparse_rows`` / ``row_elems``: the leading
    row-sparse region (needs ``ps.bitmaps``); ``zero_grads``: leave every gradient buffer cleared."""
    lib = _lib.load()
    from . import ops
    ops.bump_weights_epoch()
    for t, name in ((m, "m"), (v, "v")):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() < n:
            raise RuntimeError(f"exchange_adam_step: {name} must be a contiguous fp32 CUDA tensor of >= {n} elements")
    if grad_scale is None:
        grad_scale = 1.0 / ps.world
    if stream is None:
        stream = torch.cuda.current_stream(m.device).cuda_stream
    _lib.check(lib.nrl_exchange_adam_step(C.byref(ps), m.data_ptr(), v.data_ptr(), n, lr, beta1, beta2, eps, step,
                                          int(epoch if epoch is not None else step), grad_scale, int(max_ctas),
                                          int(timeout_s * 1e9), int(sparse_rows), int(row_elems), int(zero_grads),
                                          stream), "nrl_exchange_adam_step")
