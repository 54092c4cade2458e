"""world_size-2 tests of the N>1 host logic on CPU (gloo): the flat parameter layout, the one
exchange step of the path (gradient all-reduce + 1/world averaging folded into Adam) and the
per-rank dropout seeds.  The arithmetic on each rank is done by the oracle (this is a test of
the plumbing around the kernels, which is device-agnostic)."""
import os
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
