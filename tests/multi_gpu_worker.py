"""Worker of tests/test_gpu_peer_exchange.py::test_two_gpu_processes (launched with torchrun, one rank per GPU).

Trains the NRMS step for a few steps twice from the same initial parameters and the same per-rank batches and
dropout seeds: once with the NCCL all-reduce + dense Adam exchange, once with the fused peer-memory kernel
(``nrl_exchange_adam_step``).  Checks: barriers completed, replicas bit-identical across ranks in peer mode, and
the two modes agree (they sum the same gradients, in possibly different order)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    from newsreclib_b200.synthetic import make_batch, make_nrms_params
    from newsreclib_b200.trainer import NRMSTrainer

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    V, B, HIST = 3000, 8, 12
    params = make_nrms_params(V, 300, 15, 200, seed=11)
    batches = []
    for i in range(3):
        hb = make_batch(B, V, hist="fixed", max_hist=HIST, cand="train", seed=77 + 10 * rank + i, max_title_len=30)
        batches.append({"x_hist": {"title": hb["x_hist"]["title"].to(dev)}, "x_cand": {"title": hb["x_cand"]["title"].to(dev)},
                        "batch_hist": hb["batch_hist"].to(dev), "batch_cand": hb["batch_cand"].to(dev),
                        "labels": hb["labels"].to(dev)})
    results, init = {}, None
    for mode in ("nccl", "peer"):
        tr = NRMSTrainer(params, 15, device=dev, dropout_p=0.2, lr=1e-3, seed=5, exchange=mode)
        assert tr.world == world
        if init is None:
            init = tr.flat.clone()
        for i in range(6):
            tr.train_step(batches[i % 3], B, HIST, 5)
        torch.cuda.synchronize()
        if mode == "peer":
            assert tr.peer_block.status() == 0, f"rank {rank}: barrier timeout code {tr.peer_block.status()}"
            assert not bool(tr.grad.any()), f"rank {rank}: the exchange did not leave the gradient buffer cleared"
        flat = tr.flat.clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        for r in range(1, world):
            same = torch.equal(gathered[0], gathered[r])
            if mode == "peer":
                assert same, f"peer mode: replica {r} differs from replica 0"
        results[mode] = flat
        if mode == "peer":
            dist.barrier()
            tr.peer_block.close()
    d = (results["nccl"] - results["peer"]).abs()
    assert (results["nccl"] - init).abs().max().item() > 1e-3, "the parameters did not move"
    # six Adam steps with lr 1e-3 move a parameter by <= 6e-3.  The two modes sum the same gradients in a different
    # order; Adam divides by sqrt(v), so an element whose gradient is pure rounding noise moves by +-lr in either run
    # (by construction the key third of both in_proj_bias vectors, ~600 elements: softmax is invariant to a constant
    # key shift).  All but a small fraction of the elements must agree closely, the median difference must be tiny.
    frac, diff = (d > 2e-5).float().mean().item(), d.median().item()
    assert frac < 2e-3 and diff < 1e-6, f"rank {rank}: nccl vs peer: median |diff| {diff}, fraction off {frac}"
    dist.barrier()
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} median |nccl - peer| = {diff:.3e}, fraction > 2e-5: {frac:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
