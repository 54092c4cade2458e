"""world_size-2 tests of the N>1 host logic on CPU (gloo): the flat parameter layout, the one
exchange step of the path (gradient all-reduce + 1/world averaging folded into Adam) and the
per-rank dropout seeds.  The arithmetic on each rank is done by the oracle (this is a test of
the plumbing around the kernels, which is device-agnostic)."""
import os

import pytest
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import oracle_run
from newsreclib_b200.synthetic import make_batch, make_nrms_params
from newsreclib_b200.trainer import TITLE, USER, FlatParams, GradExchange
from newsreclib_b200.ops import BLOCK_KEYS

KEYS = [TITLE + "embedding_layer.weight"] + [TITLE + k for k in BLOCK_KEYS] + [USER + k for k in BLOCK_KEYS]
V, E, H, Q = 400, 60, 3, 40
# Adam divides by sqrt(v): with the default eps=1e-8 a gradient that is ~0 turns fp32 summation-order noise
# into a full +-lr step, so the comparison uses a large eps (the plumbing under test does not depend on it)
EPS = 1e-2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    from oracle import nrms_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    params = make_nrms_params(V, E, H, Q, seed=3)               # same replica on every rank
    fp = FlatParams(params, KEYS, "cpu")
    ex = GradExchange()
    assert ex.world == world
    for step in range(1, 3):
        batch = make_batch(4, V, hist="ragged", max_hist=5, seed=50 + 10 * step + rank)  # rank-local impressions
        _, _, g = oracle_run(fp.state_dict(), batch, H)
        fp.grad.zero_()
        for k in KEYS:
            fp.grads[k].copy_(g[k])
        scale = ex.all_reduce(fp.grad)
        O.adam_step(fp.flat, fp.grad * scale, fp.m, fp.v, step, eps=EPS)
    out[rank] = fp.flat.clone()
    dist.destroy_process_group()


def test_two_rank_exchange_matches_single_process_mean_gradient():
    from oracle import nrms_oracle as O
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert torch.equal(out[0], out[1])                         # replicas stay bit-identical
    # single process: Adam on the mean of the two ranks' gradients
    params = make_nrms_params(V, E, H, Q, seed=3)
    fp = FlatParams(params, KEYS, "cpu")
    for step in range(1, 3):
        fp.grad.zero_()
        for rank in range(world):
            batch = make_batch(4, V, hist="ragged", max_hist=5, seed=50 + 10 * step + rank)
            _, _, g = oracle_run(fp.state_dict(), batch, H)
            for k in KEYS:
                fp.grads[k].add_(g[k])
        O.adam_step(fp.flat, fp.grad * 0.5, fp.m, fp.v, step, eps=EPS)
    assert torch.allclose(out[0], fp.flat, rtol=0, atol=1e-6)


def _chunk_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.arange(1003, dtype=torch.float32) * (rank + 1)
    ex = GradExchange()
    seen = []
    for sl, work in ex.all_reduce_chunks(g, chunk_elems=256):   # 4 chunks, the last one ragged
        assert work is not None and sl.start % 4 == 0
        seen.append((sl.start, sl.stop))
    out[rank] = (g.clone(), seen)
    dist.destroy_process_group()


def test_chunked_exchange_covers_the_buffer_in_order():
    """all_reduce_chunks (Adam of chunk i overlaps the all-reduce of chunks i+1..) reduces every element once."""
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_chunk_worker, args=(world, port, out), nprocs=world, join=True)
    g0, seen = out[0]
    assert torch.equal(g0, torch.arange(1003, dtype=torch.float32) * 3) and torch.equal(g0, out[1][0])
    assert seen == [(0, 256), (256, 512), (512, 768), (768, 1003)]
    one = list(GradExchange().all_reduce_chunks(torch.zeros(10)))  # world 1: a single slice, no communication
    assert len(one) == 1 and one[0][1] is None and (one[0][0].start, one[0][0].stop) == (0, 10)


def test_flat_layout_and_seeds():
    params = make_nrms_params(V, E, H, Q, seed=1)
    fp = FlatParams(params, KEYS, "cpu")
    assert all(o % 4 == 0 for o in fp.offsets)                 # 16-byte aligned starts
    sd = fp.state_dict()
    assert list(sd) == KEYS and all(torch.equal(sd[k], params[k]) for k in KEYS)
    fp.flat.add_(1.0)                                          # views alias the flat buffer
    assert torch.equal(fp.params[KEYS[3]], params[KEYS[3]] + 1.0)
    assert GradExchange().world == 1 and GradExchange().all_reduce(fp.grad) == 1.0
    seeds = {GradExchange.rank_seed(1234, r, s) for r in range(8) for s in range(1000)}
    assert len(seeds) == 8000


def _sharded_worker(rank, world, port, out):
    """CPU restatement of what nrl_exchange_adam_step does on a rank (csrc/nrl_exchange.cuh): rank-ordered sum of the
    ranks' gradients on the OWNED slice, Adam there, new parameters to every replica; moments stay sharded."""
    from newsreclib_b200.exchange import slice_bounds
    from oracle import nrms_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 4 * 257
    gen = torch.Generator().manual_seed(7)
    p = torch.randn(n, generator=gen)
    m, v = torch.zeros(n), torch.zeros(n)
    lo, hi = slice_bounds(n, world, rank)
    for step in (1, 2, 3):
        grads = [torch.randn(n, generator=gen) * (r + 1) for r in range(world)]   # every rank can see every gradient here
        g = grads[0][lo:hi].clone()
        for r in range(1, world):
            g += grads[r][lo:hi]
        O.adam_step(p[lo:hi], g * (1.0 / world), m[lo:hi], v[lo:hi], step, eps=EPS)
        new = torch.zeros(n)
        new[lo:hi] = p[lo:hi]
        dist.all_reduce(new)                                  # "peer stores": every replica receives every owned slice
        p.copy_(new)
    ex = GradExchange()
    out[rank] = (p.clone(), ex.gather_sharded(m), ex.gather_sharded(v), m.clone())
    dist.destroy_process_group()


def test_sharded_exchange_equals_all_reduce_plus_dense_adam():
    from newsreclib_b200.exchange import slice_bounds
    from oracle import nrms_oracle as O
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_sharded_worker, args=(world, port, out), nprocs=world, join=True)
    n = 4 * 257
    gen = torch.Generator().manual_seed(7)
    p = torch.randn(n, generator=gen)
    m, v = torch.zeros(n), torch.zeros(n)
    for step in (1, 2, 3):
        grads = [torch.randn(n, generator=gen) * (r + 1) for r in range(world)]
        O.adam_step(p, (grads[0] + grads[1]) * 0.5, m, v, step, eps=EPS)
    for rank in range(world):
        pr, mr, vr, m_local = out[rank]
        assert torch.equal(pr, p) and torch.equal(mr, m) and torch.equal(vr, v)   # same elementwise arithmetic: same bits
        lo, hi = slice_bounds(n, world, rank)
        assert not bool(m_local[:lo].any()) and not bool(m_local[hi:].any()) and bool(m_local[lo:hi].any())


def test_slice_ownership_partitions_the_flat_buffer():
    from newsreclib_b200.exchange import slice_bounds
    for n in (4, 8, 12, 4 * 1001, 4 * 21_843_200 // 4):
        for world in (1, 2, 3, 4, 5, 8, 16):
            edges = [slice_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))            # contiguous, disjoint, in rank order
            assert all(lo % 4 == 0 and hi % 4 == 0 and lo <= hi for lo, hi in edges)  # 16-byte granules
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(s for s in sizes if s) <= 4 * ((n // 4 + world - 1) // world)  # at most one granule row
    with pytest.raises(ValueError):
        slice_bounds(10, 2, 0)


class _FakeLib:
    """Stands in for the C ABI's peer-memory calls: hands out fake base addresses, optionally fails one call."""

    def __init__(self, rank, fail_alloc_on=None, fail_open_on=None):
        self.rank, self.fail_alloc_on, self.fail_open_on = rank, fail_alloc_on, fail_open_on
        self.closed, self.freed = [], []

    def nrl_peer_alloc(self, nbytes, base_ref, handle):
        if self.rank == self.fail_alloc_on:
            return -3
        base_ref._obj.value = 0x10000000 * (self.rank + 1)
        handle.raw = bytes([self.rank + 1]) * 64
        return 0

    def nrl_peer_open(self, handle, ptr_ref):
        if self.rank == self.fail_open_on:
            return -3
        ptr_ref._obj.value = 0x10000000 * handle[0] + 0x1000        # "mapped" address of the exporting rank's block
        return 0

    def nrl_peer_close(self, ptr):
        self.closed.append(ptr)
        return 0

    def nrl_peer_free(self, ptr):
        self.freed.append(ptr)
        return 0


def _peer_block_worker(rank, world, port, out, fail_alloc_on, fail_open_on):
    import contextlib
    from newsreclib_b200.exchange import PeerBlock
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class HostPeerBlock(PeerBlock):          # the two seams: no CUDA device context, no aliasing of fake addresses
        def _device_ctx(self):
            return contextlib.nullcontext()

        def _alias(self, ptr, nelem):
            return torch.zeros(nelem)

    lib = _FakeLib(rank, fail_alloc_on, fail_open_on)
    n = 64
    try:
        pb = HostPeerBlock(n, torch.device("cpu"), lib=lib)
        ps = pb.peer_set
        out[rank] = ("ok", [int(ps.params[r] or 0) for r in range(world)], [int(ps.grads[r] or 0) for r in range(world)],
                     [int(ps.flags[r] or 0) for r in range(world)], int(ps.world), int(ps.rank))
        dist.barrier()
        pb.close()
        assert lib.freed == [0x10000000 * (rank + 1)] and len(lib.closed) == world - 1
    except RuntimeError as e:
        out[rank] = ("error", str(e), list(lib.freed), list(lib.closed))
    dist.barrier()                           # nobody is left behind in a collective, whichever rank failed
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_alloc_on,fail_open_on", [(None, None), (1, None), (None, 0)])
def test_peer_block_construction_protocol(fail_alloc_on, fail_open_on):
    """PeerBlock's handle exchange on two gloo ranks with the C ABI's peer calls faked: on success every rank holds
    every rank's block addresses in the layout [params | grads | flags]; when ONE rank cannot allocate or map, ALL ranks
    raise the same error (no rank is left waiting in a collective) and what was allocated / mapped is released."""
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_peer_block_worker, args=(world, port, out, fail_alloc_on, fail_open_on), nprocs=world, join=True)
    n = 64
    if fail_alloc_on is None and fail_open_on is None:
        for rank in range(world):
            status, params, grads, flags, w, r = out[rank]
            assert status == "ok" and (w, r) == (world, rank)
            for peer in range(world):
                base = 0x10000000 * (peer + 1) + (0 if peer == rank else 0x1000)
                assert (params[peer], grads[peer], flags[peer]) == (base, base + 4 * n, base + 8 * n)
    else:
        bad = fail_alloc_on if fail_alloc_on is not None else fail_open_on
        for rank in range(world):
            status, msg, freed, closed = out[rank]
            assert status == "error" and "peer exchange unavailable" in msg and f"rank {bad}" in msg
            if fail_alloc_on is None:
                assert freed == [0x10000000 * (rank + 1)]                   # own block released on every rank
                assert len(closed) == (0 if rank == bad else world - 1)      # the healthy rank unmaps what it mapped
            else:
                assert freed == ([] if rank == bad else [0x10000000 * (rank + 1)]) and closed == []


def _make_trainer_worker(rank, world, port, scenario, out):
    """bench.make_trainer(--exchange auto) with a fake trainer class: which exchange do the ranks end up with when the
    peer path is unavailable on ONE rank, times out in the probe on one rank, or leaves different replicas?"""
    import bench
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakePeer:
        def __init__(self, status):
            self._status, self.closed = status, False

        def status(self):
            return self._status

        def close(self):
            self.closed = True

    class FakeTrainer:
        def __init__(self, params, H, device=None, dropout_p=0.0, precision=0, exchange="nccl", exchange_timeout_s=30.0):
            self.mode = exchange
            if exchange == "peer" and scenario == "ctor_fails_on_rank1" and rank == 1:
                raise RuntimeError("peer exchange unavailable: rank 1: no IPC")
            self.flat, self.m, self.v, self.step_count = torch.zeros(16), torch.ones(16), torch.ones(16), 0
            self.peer_block = None
            if exchange == "peer":
                self.peer_block = FakePeer(1 if (scenario == "probe_times_out_on_rank0" and rank == 0) else 0)

        def train_step(self, b, B, Hmax, Cmax):
            self.step_count += 1
            self.flat += 2.0 if (scenario == "replicas_differ" and rank == 1) else 1.0

        def probe_exchange(self, b, B, Hmax, Cmax):
            self.train_step(b, B, Hmax, Cmax)
            return True, True

    tr, used = bench.make_trainer(FakeTrainer, {}, torch.device("cpu"), 0, world, "auto")
    out[rank] = (tr.mode, used)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scenario", ["ctor_fails_on_rank1", "probe_times_out_on_rank0", "replicas_differ"])
def test_bench_auto_exchange_falls_back_on_every_rank(scenario):
    """N > 1, --exchange auto: if the fused peer exchange cannot be set up or fails its probe on ANY rank, EVERY rank
    must end up on the NCCL exchange (and say why) -- a split decision would deadlock the first step."""
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_make_trainer_worker, args=(world, port, scenario, out), nprocs=world, join=True)
    for rank in range(world):
        mode, used = out[rank]
        assert mode == "nccl" and used.startswith("nccl all-reduce + dense Adam (peer exchange unavailable")
