#!/usr/bin/env python
"""Headline benchmark: NRMS training-step throughput (impressions/s) on synthetic MIND-shaped
batches, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch: forward + soft-target CE + backward + Adam
(+ the gradient all-reduce when N > 1), train mode with dropout 0.2 — BASELINE.json configs[1]
(NRMS synthetic MINDsmall-shape: 300-d embeddings, 15 heads, 50-news histories, 30-token
titles, batch 64 per GPU, 1 positive + 4 negatives, V = 70 000).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from newsreclib_b200.synthetic import make_batch, make_nrms_params  # noqa: E402

E, H, Q, L, HIST, CAND, VOCAB, DROPOUT = 300, 15, 200, 30, 50, 5, 70000, 0.2
WORKLOAD = ("NRMS train step (fwd + soft-target CE + bwd + Adam), MINDsmall-shape synthetic: "
            "B=64/GPU, hist 50 (fixed), 5 candidates, 30-token titles, E=300, 15 heads, Q=200, "
            "V=70000, dropout 0.2")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def gemm_traffic_per_launch():
    """DRAM bytes per GEMM launch from the committed ncu --set full capture (None if absent)."""
    p = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    try:
        return json.load(open(p))["dram_bytes_per_launch_avg"]
    except (OSError, KeyError, ValueError):
        return None


# ----------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def window(self, t0, t1):
        """Host-clock window of the timed region; only samples inside it are reported."""
        self.t0, self.t1 = t0, t1

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 <= t <= self.t1 + 0.03)]
        if not rows:
            rows = [r for t, r in self.rows[-3:]]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference on the host cores (baseline/_ref, installed by baseline/install_ref.py);
# the oracle port only if that install is absent
# ----------------------------------------------------------------------------------------
def run_config(world: int, batch: int):
    """The `config` object of the JSON line -- the same in both arms (the workload, not the implementation)."""
    return {"workload": WORKLOAD, "global_batch": world * batch, "batch_per_gpu": batch, "parallelism": f"dp{world}",
            "l2": "inputs rotate over 4 distinct batches; each step streams ~2 GB of activations "
                  "(>> 126 MB L2), so no step starts with a warm L2"}


def _reference_stepper(threads: int):
    """One train step of the reference's own code: `NRMSModule.model_step` (nrms_module.py:260-362: forward :230-255,
    to_dense_batch, CrossEntropyLoss, the per-row output collection) -> `loss.backward()` -> the optimizer built by the
    reference's `configure_optimizers` (abstract_recommender.py:89-108) from `torch.optim.Adam(lr=1e-4)`
    (configs/model/nrms.yaml:49-52), train mode (dropout 0.2), fp32.  The module files are loaded unmodified from
    baseline/_ref under the stand-ins of oracle/ref_standins.py (lightning / torchmetrics / torch_geometric are absent;
    `to_dense_batch` is the restated PyG 2.3.0 function)."""
    import functools
    import tempfile
    import numpy as np
    from oracle import ref_standins

    ref_standins.install(ref_standins.BASELINE_REF if os.path.isdir(os.path.join(ref_standins.BASELINE_REF, "newsreclib"))
                         else None)
    from newsreclib.models.general_rec.nrms_module import NRMSModule

    params = make_nrms_params(VOCAB, E, H, Q, seed=1234)
    outputs = {k: ["preds", "targets", "cand_news_size"] for k in ("train", "val", "test")}
    with tempfile.TemporaryDirectory() as tmp:
        emb = os.path.join(tmp, "emb.npy")
        np.save(emb, params["news_encoder.text_encoders.title.embedding_layer.weight"].numpy())
        m = NRMSModule(
            dataset_attributes=["title", "category"], attributes2encode=["title"], outputs=outputs,
            dual_loss_training=False, dual_loss_coef=None, loss="cross_entropy_loss", late_fusion=False,
            temperature=None, use_plm=False, pretrained_embeddings_path=emb, plm_model=None, frozen_layers=None,
            embed_dim=E, num_heads=H, query_dim=Q, dropout_probability=DROPOUT, top_k_list=[5, 10],
            num_categ_classes=18, num_sent_classes=3, save_recs=False, recs_fpath=None,
            optimizer=functools.partial(torch.optim.Adam, lr=1e-4), scheduler=None)
    res = m.load_state_dict({k: v for k, v in params.items() if k in m.state_dict()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    m.train()
    opt = m.configure_optimizers()["optimizer"]

    def one(bs, seed):
        batch = make_batch(bs, VOCAB, hist="fixed", max_hist=HIST, cand="train", seed=seed, max_title_len=L)
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = m.model_step(batch)[0]
        loss.backward()
        opt.step()
        return time.perf_counter() - t0
    return one, "reference"


def _port_stepper(threads: int):
    """Fallback when baseline/_ref is absent: the oracle port of the same modules (oracle/nrms_oracle.py), train mode
    with fresh Bernoulli masks, CE, autograd backward, torch.optim.Adam over all parameters incl. the dense table."""
    from oracle import nrms_oracle as O

    params = {k: v.clone().requires_grad_(True) for k, v in make_nrms_params(VOCAB, E, H, Q, seed=1234).items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)

    def one(bs, seed):
        batch = make_batch(bs, VOCAB, hist="fixed", max_hist=HIST, cand="train", seed=seed, max_title_len=L)
        nh, nc = batch["x_hist"]["title"].shape[0], batch["x_cand"]["title"].shape[0]
        t0 = time.perf_counter()
        keep = 1.0 - DROPOUT
        masks = {"hist1": torch.empty(nh, L, E).bernoulli_(keep), "hist2": torch.empty(nh, L, E).bernoulli_(keep),
                 "cand1": torch.empty(nc, L, E).bernoulli_(keep), "cand2": torch.empty(nc, L, E).bernoulli_(keep)}
        opt.zero_grad(set_to_none=True)
        scores = O.nrms_forward(batch, params, H, masks=masks, dropout_p=DROPOUT)
        loss = O.nrms_loss(batch, scores)
        loss.backward()
        params["news_encoder.text_encoders.title.embedding_layer.weight"].grad[0] = 0  # padding_idx
        opt.step()
        return time.perf_counter() - t0
    return one, "port"


def cpu_train_steps(batch_size: int, steps: int, warmup: int, budget_s: float):
    """The reference's CPU path on all host threads; returns (impressions per step, timed step seconds, threads,
    kind).  The sample is bounded: if `steps + warmup` full batches would exceed `budget_s`, every step processes
    fewer impressions of the same shape."""
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    try:
        one, kind = _reference_stepper(threads)
    except ImportError as e:
        print(f"[bench] reference install unavailable ({e}); timing the oracle port instead", file=sys.stderr)
        one, kind = _port_stepper(threads)
    bs = batch_size
    t_first = one(bs, 0)
    total = steps + warmup
    if t_first * total > budget_s:  # bound the sample: fewer impressions per step, same shape
        bs = max(4, int(batch_size * budget_s / (t_first * total)))
    times = [one(bs, 1 + i) for i in range(total)]
    timed = times[warmup:] if len(times) > warmup else times
    return bs, timed, threads, kind


CPU_KIND_TEXT = {"reference": "the reference's own NRMSModule.model_step + backward + its configure_optimizers Adam "
                              "(unmodified files from baseline/_ref)",
                 "port": "oracle port of the reference modules"}


def run_reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    bs, timed, threads, kind = cpu_train_steps(args.batch, args.steps, args.warmup, budget_s=150.0)
    ms = 1e3 * sum(timed) / len(timed)
    val = bs / (ms / 1e3)
    sample = (f"{len(timed)} train steps of {bs} impressions (same shape as the GPU workload) on the host CPU, fp32, "
              f"{threads} threads: {CPU_KIND_TEXT[kind]}")
    print(json.dumps({
        "impl": "reference", "metric": "impressions/sec", "value": val, "unit": "impressions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": run_config(world, args.batch),
        "cpu_baseline": {"value": val, "unit": "impressions/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "impressions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
GEMM_FLOPS = {  # algorithmic (unpadded, single-pass) flops per row of the block
    "gemm in_proj": lambda: 2 * E * 3 * E, "gemm out_proj": lambda: 2 * E * E, "gemm additive": lambda: 2 * E * Q,
    "gemm additive dgrad": lambda: 2 * E * Q, "gemm additive wgrad": lambda: 2 * E * Q,
    "gemm out_proj dgrad": lambda: 2 * E * E, "gemm out_proj wgrad": lambda: 2 * E * E,
    "gemm in_proj dgrad": lambda: 2 * E * 3 * E, "gemm in_proj wgrad": lambda: 2 * E * 3 * E,
}


def main_naml(args, rank, local_rank, world):
    """Secondary line: NAML (title + abstract + category views, CNN text encoder; BASELINE.json configs[4]
    shape at MINDsmall vocabulary) trained through the drop-in NAMLModule + ModuleTrainer."""
    import functools
    import torch.distributed as dist
    from newsreclib_b200 import _lib
    from newsreclib_b200.models.general_rec.naml_module import NAMLModule
    from newsreclib_b200.synthetic import make_naml_params
    from newsreclib_b200.trainer import ModuleTrainer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: newsreclib_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B, F_, W, CE, LA = args.batch, 400, 3, 100, 50
    params = make_naml_params(VOCAB, E, F_, W, Q, CE, 19, seed=1234)
    outputs = {k: ["preds", "targets", "cand_news_size"] for k in ("train", "val", "test")}
    m = NAMLModule(
        dataset_attributes=["title", "abstract", "category", "subcategory"], attributes2encode=["title", "abstract", "category"],
        outputs=outputs, dual_loss_training=False, dual_loss_coef=None, loss="cross_entropy_loss", late_fusion=False,
        temperature=None, use_plm=False, pretrained_embeddings_path=None, plm_model=None, frozen_layers=None,
        text_embed_dim=E, num_heads=H, num_filters=F_, window_size=W, query_dim=Q, categ_embed_dim=CE,
        dropout_probability=DROPOUT, top_k_list=[5, 10], num_categ_classes=18, num_sent_classes=3, save_recs=False,
        recs_fpath=None, optimizer=functools.partial(torch.optim.Adam, lr=1e-4), scheduler=None,
        pretrained_embeddings=params["news_encoder.text_encoders.title.embedding_layer.weight"])
    full = dict(params)
    for k in list(params):
        if ".text_encoders.title." in k:
            full[k.replace(".title.", ".abstract.")] = params[k]
    m.load_state_dict(full)
    m = m.to(dev)
    tr = ModuleTrainer(m, lr=1e-4)

    def to_dev(hb):
        return {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else
                    {kk: vv.to(dev, non_blocking=True) for kk, vv in v.items()} if isinstance(v, dict) else v)
                for k, v in hb.items()}
    host = [make_batch(B, VOCAB, hist="fixed", max_hist=HIST, cand="train", seed=1234 + rank * 100 + i, max_title_len=L,
                       abstract_len=LA) for i in range(4)]
    host = [{k: (v.pin_memory() if torch.is_tensor(v) else {kk: vv.pin_memory() for kk, vv in v.items()} if isinstance(v, dict) else v)
             for k, v in hb.items()} for hb in host]
    devb = [to_dev(hb) for hb in host]

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)
    for i in range(args.warmup):
        tr.train_step(devb[i % 4])
    l0 = lib.nrl_launch_count()
    ms = timed(lambda i: tr.train_step(devb[i % 4]), args.steps) / args.steps
    launches = lib.nrl_launch_count() - l0
    ms_e2e = timed(lambda i: float(tr.train_step(to_dev(host[i % 4]))), args.steps) / args.steps
    if rank == 0:
        nh, nc = B * HIST, B * CAND
        h2d = sum(v.numel() * v.element_size() if torch.is_tensor(v) else sum(x.numel() * x.element_size() for x in v.values())
                  for v in host[0].values())
        print(json.dumps({
            "metric": "impressions/sec", "value": world * B / (ms / 1e3), "unit": "impressions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32 via bf16x3 split on tcgen05 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": f"NAML train step through NAMLModule + ModuleTrainer (fwd + CE + autograd bwd + Adam): B={B}/GPU, "
                                   f"hist {HIST}, {CAND} candidates, title {L} + abstract {LA} tokens + category, E={E}, "
                                   f"F={F_}, w={W}, Q={Q}, V={VOCAB}, dropout {DROPOUT}", "global_batch": world * B,
                       "parallelism": f"dp{world}", "news_per_step": nh + nc},
            "e2e": {"value": world * B / (ms_e2e / 1e3), "unit": "impressions/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------
# the other BASELINE.json configurations, as extra keys of the one JSON line (the headline stays configs[1])
# ----------------------------------------------------------------------------------------
def _timed_steps(fn, steps, warmup, dev, world):
    import torch.distributed as dist
    for i in range(warmup):
        fn(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / steps


def _dev_batch(hb, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else {kk: vv.to(dev) for kk, vv in v.items()} if isinstance(v, dict) else v)
            for k, v in hb.items()}


def _module_kwargs(outputs):
    import functools
    return dict(outputs=outputs, dual_loss_training=False, dual_loss_coef=None, loss="cross_entropy_loss", late_fusion=False,
                temperature=None, dropout_probability=DROPOUT, top_k_list=[5, 10], num_categ_classes=18, num_sent_classes=3,
                save_recs=False, recs_fpath=None, optimizer=functools.partial(torch.optim.Adam, lr=1e-4), scheduler=None)


def extra_configs(args, dev, rank, world, exchange_mode, steps=15, warmup=3):
    """Secondary measurements, each a short timed loop of full training steps on synthetic data of the named shape
    (device-resident inputs, max over ranks).  Every entry is independent and guarded: a failure is reported in place and
    never costs the headline."""
    from newsreclib_b200 import ops
    from newsreclib_b200.trainer import ModuleTrainer, NRMSTrainer
    B = args.batch
    peer = exchange_mode.startswith("peer")
    outputs = {k: ["preds", "targets", "cand_news_size"] for k in ("train", "val", "test")}
    out = {}

    def entry(name, fn):
        try:
            torch.cuda.synchronize()
            out[name] = fn()
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    def imps(ms, per_gpu):
        return {"value": world * per_gpu / (ms / 1e3), "unit": "impressions/s", "ms_per_step": ms}

    # -- configs[2]: NRMS MINDlarge-shape (V = 130 000), single-pass bf16, the fused step
    def nrms_large_bf16():
        V = 130000
        tr = NRMSTrainer(make_nrms_params(V, E, H, Q, seed=1234), H, device=dev, dropout_p=DROPOUT, precision=ops.PREC_BF16,
                         exchange="peer" if peer else "nccl", exchange_timeout_s=5.0, status_every=0)
        bs = [_dev_batch(make_batch(B, V, hist="fixed", max_hist=HIST, cand="train", seed=500 + rank * 100 + i, max_title_len=L), dev)
              for i in range(4)]
        bs = [{"x_hist": {"title": b["x_hist"]["title"]}, "x_cand": {"title": b["x_cand"]["title"]}, "batch_hist": b["batch_hist"],
               "batch_cand": b["batch_cand"], "labels": b["labels"]} for b in bs]
        ms = _timed_steps(lambda i: tr.train_step(bs[i % 4], B, HIST, CAND), steps, warmup, dev, world)
        tr.check_status()
        r = imps(ms, B)
        r.update(workload=f"BASELINE configs[2]: NRMS train step, MINDlarge-shape V={V}, B={B}/GPU, hist {HIST}, {CAND} candidates, "
                          f"single-pass bf16 (fp32 accumulate), {'peer exchange' if peer else 'no exchange' if world == 1 else 'nccl'}")
        if tr.peer_block is not None:
            torch.distributed.barrier()
            tr.peer_block.close()
        return r
    entry("nrms_mindlarge_bf16", nrms_large_bf16)

    # -- configs[4]: NAML (title + abstract + category), MINDlarge-shape, through the drop-in NAMLModule
    def naml():
        from newsreclib_b200.models.general_rec.naml_module import NAMLModule
        from newsreclib_b200.synthetic import make_naml_params
        V, F_, W, CE, LA = 130000, 400, 3, 100, 50
        params = make_naml_params(V, E, F_, W, Q, CE, 19, seed=1234)
        m = NAMLModule(dataset_attributes=["title", "abstract", "category", "subcategory"],
                       attributes2encode=["title", "abstract", "category"], use_plm=False, pretrained_embeddings_path=None,
                       plm_model=None, frozen_layers=None, text_embed_dim=E, num_heads=H, num_filters=F_, window_size=W,
                       query_dim=Q, categ_embed_dim=CE,
                       pretrained_embeddings=params["news_encoder.text_encoders.title.embedding_layer.weight"],
                       **_module_kwargs(outputs))
        full = dict(params)
        for k in list(params):
            if ".text_encoders.title." in k:
                full[k.replace(".title.", ".abstract.")] = params[k]
        m.load_state_dict(full)
        tr = ModuleTrainer(m.to(dev), lr=1e-4, exchange="peer" if peer else "nccl", exchange_timeout_s=5.0)
        bs = [_dev_batch(make_batch(B, V, hist="fixed", max_hist=HIST, cand="train", seed=700 + rank * 100 + i, max_title_len=L,
                                    abstract_len=LA), dev) for i in range(4)]
        ms = _timed_steps(lambda i: tr.train_step(bs[i % 4]), steps, warmup, dev, world)
        tr.check_status()
        r = imps(ms, B)
        r.update(workload=f"BASELINE configs[4]: NAML train step through NAMLModule + ModuleTrainer (autograd over the sm_100a "
                          f"ops), MINDlarge-shape V={V}, B={B}/GPU, title {L} + abstract {LA} tokens + category, F={F_}, w={W}")
        if tr.peer_block is not None:
            torch.distributed.barrier()
            tr.peer_block.close()
        return r
    entry("naml_mindlarge", naml)

    # -- the drop-in path a Lightning user gets: NRMSModule.model_step + autograd + an optimizer
    def nrms_module():
        from newsreclib_b200.models.general_rec.nrms_module import NRMSModule
        params = make_nrms_params(VOCAB, E, H, Q, seed=1234)

        def build():
            m = NRMSModule(dataset_attributes=["title", "category"], attributes2encode=["title"], use_plm=False,
                           pretrained_embeddings_path=None, plm_model=None, frozen_layers=None, embed_dim=E, num_heads=H,
                           query_dim=Q, pretrained_embeddings=params["news_encoder.text_encoders.title.embedding_layer.weight"],
                           **_module_kwargs(outputs))
            m.load_state_dict({k: v for k, v in params.items() if k in m.state_dict()})
            return m.to(dev).train()
        bs = [_dev_batch(make_batch(B, VOCAB, hist="fixed", max_hist=HIST, cand="train", seed=900 + rank * 100 + i,
                                    max_title_len=L), dev) for i in range(4)]
        tr = ModuleTrainer(build(), lr=1e-4, exchange="peer" if peer else "nccl", exchange_timeout_s=5.0)
        ms = _timed_steps(lambda i: tr.train_step(bs[i % 4]), steps, warmup, dev, world)
        tr.check_status()
        r = imps(ms, B)
        r.update(workload="NRMSModule.model_step (one autograd node on nrl_nrms_step / nrl_nrms_step_bwd) + loss.backward() + "
                          "ModuleTrainer (flat fused Adam, gradients accumulated in place): the drop-in module path, headline "
                          "workload")
        if tr.peer_block is not None:
            torch.distributed.barrier()
            tr.peer_block.close()
        if world == 1:  # the same module on the per-op autograd.Functions (ten nodes per step), as before the fused node
            m = build()
            m.fused_model_step = False
            tr2 = ModuleTrainer(m, lr=1e-4, exchange="nccl")
            r["per_op_functions"] = imps(_timed_steps(lambda i: tr2.train_step(bs[i % 4]), steps, warmup, dev, world), B)
            del tr2, m
        if world == 1:  # stock torch.optim.Adam from the module's own configure_optimizers (what Lightning would call)
            m = build()
            opt = m.configure_optimizers()["optimizer"]

            def step(i):
                opt.zero_grad(set_to_none=True)
                m.training_step(bs[i % 4], i).backward()
                opt.step()
            ms2 = _timed_steps(step, steps, warmup, dev, world)
            r["with_torch_optim_adam"] = imps(ms2, B)
            # the same loop with optimizer._target_ = newsreclib_b200.optim.Adam (one Adam launch per parameter)
            from newsreclib_b200.optim import Adam as NrlAdam
            del m, opt
            m = build()
            opt = NrlAdam(m.parameters(), lr=1e-4)
            r["with_nrl_optim_adam"] = imps(_timed_steps(step, steps, warmup, dev, world), B)
        return r
    entry("nrms_module_dropin", nrms_module)

    # -- configs[3]: NRMS-PLM, roberta-base-shaped news encoder (random init), layers 0-7 frozen, B = 8 per GPU.
    # transformer_impl="native": embeddings + 12 layers on the sm_100a encoder + the MHSA / additive
    # head; the HF torch module on the same weights and batches beside it ("with_hf_transformer").
    def nrms_plm():
        import numpy as np
        from transformers import RobertaConfig, RobertaModel
        from newsreclib_b200.models.general_rec.nrms_module import NRMSModule
        Bp, Ep, Hp = 8, 768, 16

        def run(impl, max_len, precision=None):
            torch.manual_seed(1234)
            plm = RobertaModel(RobertaConfig(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1),
                               add_pooling_layer=False)
            m = NRMSModule(dataset_attributes=["title", "category"], attributes2encode=["title"], use_plm=True,
                           pretrained_embeddings_path=None, plm_model=plm, frozen_layers=list(range(8)), embed_dim=Ep,
                           num_heads=Hp, query_dim=Q, transformer_impl=impl, **_module_kwargs(outputs))
            if precision is not None:
                m.news_encoder.text_encoders["title"].precision = precision
            tr = ModuleTrainer(m.to(dev), lr=1e-5, exchange="nccl")
            rng = np.random.default_rng(1234 + rank)

            def plm_news(n):
                lens = np.clip(rng.poisson(16, n), 6, max_len)
                lens[0] = max_len if max_len == 96 else lens[0]  # tokenizer(padding=True): the batch is padded to its longest
                T = int(lens.max())
                ids = rng.integers(3, 50265, (n, T))
                mask = np.arange(T)[None, :] < lens[:, None]
                ids[~mask] = 1
                return {"input_ids": torch.from_numpy(ids).to(dev), "attention_mask": torch.from_numpy(mask.astype(np.int64)).to(dev)}
            bs = []
            for i in range(2):
                hb = make_batch(Bp, 1000, hist="fixed", max_hist=HIST, cand="train", seed=40 + rank * 10 + i, max_title_len=L)
                b = _dev_batch(hb, dev)
                b["x_hist"]["title"], b["x_cand"]["title"] = plm_news(Bp * HIST), plm_news(Bp * CAND)
                bs.append(b)
            ms = _timed_steps(lambda i: tr.train_step(bs[i % 2]), 5, 2, dev, world)
            r = imps(ms, Bp)
            r["tokens_per_title_padded"] = [int(b["x_hist"]["title"]["input_ids"].shape[1]) for b in bs]
            if impl == "native" and rank == 0 and world == 1:
                # live per-launch times of one more step: the transformer's tcgen05 projections against the bf16 peak.
                # Algorithmic FLOPs per token row and layer: 2 D (3D + D + 2I) forward, the same for the data gradients
                # (every layer: the embeddings train), again for the weight gradients of the trainable layers 8-11.
                import ctypes as C
                from newsreclib_b200 import _lib
                lib = _lib.load()
                torch.cuda.synchronize()
                lib.nrl_profile_start(torch.cuda.current_stream().cuda_stream)
                tr.train_step(bs[0])
                torch.cuda.synchronize()
                names, msbuf = C.create_string_buffer(4000 * 48), (C.c_float * 4000)()
                n = lib.nrl_profile_stop(names, 48, msbuf, 4000)
                gemm_ms = sum(msbuf[i] for i in range(n) if names.raw[i * 48:(i + 1) * 48].split(b"\0")[0].startswith(b"tfm gemm"))
                attn_ms = sum(msbuf[i] for i in range(n) if names.raw[i * 48:(i + 1) * 48].split(b"\0")[0].startswith(b"tfm attn"))
                rows = sum(int(bs[0][k]["title"]["input_ids"].numel()) for k in ("x_hist", "x_cand"))
                per_row = 2 * 768 * (3 * 768 + 768 + 2 * 3072)
                flops = rows * per_row * (12 + 12 + 4)
                _, _, tf_sust, src = peaks()
                ach = flops / (gemm_ms / 1e3) / 1e12
                r["roofline"] = {"bound": "tensor", "kernel": "nrl_gemm_tc_kernel / nrl_gemm_tc2_kernel (the transformer's projections of one step)",
                                 "achieved": ach, "peak": tf_sust, "unit": "TFLOP/s", "frac": ach / tf_sust,
                                 "peak_source": f"{src} bf16 dense, sustained", "issued_passes": 1 if precision is not None else 3,
                                 "algorithmic_gflop_per_step": flops / 1e9, "kernel_ms_per_step": gemm_ms,
                                 "token_rows": rows, "tfm_attention_ms_per_step": attn_ms}
            del tr, m, plm, bs
            torch.cuda.empty_cache()
            return r
        r = run("native", 40)
        r.update(workload=f"BASELINE configs[3]: NRMS-PLM train step, roberta-base-shaped encoder (random init, layers 0-7 "
                          f"frozen, embeddings + layers 8-11 trained): transformer on the sm_100a encoder (bf16x3 tcgen05 GEMMs, "
                          f"tensor-core attention, fp32-equivalent) + sm_100a MHSA/additive head (768-d, 16 heads), B={Bp}/GPU, "
                          f"{HIST} + {CAND} news per impression, titles of 6-40 tokens padded to the longest of the batch, all "
                          f"three dropouts on, nccl exchange")
        r["with_hf_transformer"] = run("hf", 40)
        r["titles_padded_to_96_tokens"] = {"native": run("native", 96), "hf": run("hf", 96)}
        from newsreclib_b200 import ops as _ops
        r["native_bf16_single_pass"] = run("native", 40, _ops.PREC_BF16)
        return r
    entry("nrms_plm_roberta_base", nrms_plm)

    # -- evaluation at MINDlarge-dev shape (whole impressions, candidate lists up to 300): news
    # vectors encoded once per epoch for the news table, then an impression batch = two row gathers + user encoder +
    # scorer + device metrics; the plain (re-encoding) eval forward of the same batches beside it
    def eval_cached():
        from newsreclib_b200.metrics import ranking_metrics
        from newsreclib_b200.models.general_rec.nrms_module import NRMSModule
        from newsreclib_b200.synthetic import make_titles
        import numpy as np
        V, M = 130000, 100000  # MINDlarge: ~130 k vocabulary rows, ~100 k news in the dev table
        params = make_nrms_params(V, E, H, Q, seed=1234)
        m = NRMSModule(dataset_attributes=["title", "category"], attributes2encode=["title"], use_plm=False,
                       pretrained_embeddings_path=None, plm_model=None, frozen_layers=None, embed_dim=E, num_heads=H,
                       query_dim=Q, pretrained_embeddings=params["news_encoder.text_encoders.title.embedding_layer.weight"],
                       **_module_kwargs(outputs))
        m.load_state_dict({k: v for k, v in params.items() if k in m.state_dict()})
        m = m.to(dev).eval()
        rng = np.random.default_rng(7 + rank)
        titles = torch.from_numpy(make_titles(rng, M, V, L)).to(dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vecs = m.encode_news_table({"title": titles})
        e1.record()
        torch.cuda.synchronize()
        enc_ms = e0.elapsed_time(e1)
        bs = []
        for i in range(4):
            hb = make_batch(B, V, hist="ragged", max_hist=HIST, cand="eval", seed=300 + rank * 100 + i, max_title_len=L)
            nh, nc = hb["batch_hist"].numel(), hb["batch_cand"].numel()
            hr, cr = torch.from_numpy(rng.integers(0, M, nh)).to(dev), torch.from_numpy(rng.integers(0, M, nc)).to(dev)
            full = _dev_batch(hb, dev)
            full["x_hist"]["title"], full["x_cand"]["title"] = titles[hr], titles[cr]
            bs.append((hr, cr, full, torch.bincount(hb["batch_cand"], minlength=B).to(dev)))

        def cached(i):
            hr, cr, full, sizes = bs[i % 4]
            sc = m.forward_cached(vecs, hr, full["batch_hist"], cr, full["batch_cand"], B)
            mask = torch.arange(sc.shape[1], device=dev)[None, :] < sizes[:, None]
            return ranking_metrics(sc[mask], full["labels"], sizes, [5, 10])

        def plain(i):
            with torch.no_grad():
                return m(bs[i % 4][2])
        def cached_scores_only(i):
            hr, cr, full, sizes = bs[i % 4]
            return m.forward_cached(vecs, hr, full["batch_hist"], cr, full["batch_cand"], B)
        ms_c = _timed_steps(cached, steps, warmup, dev, world)
        ms_s = _timed_steps(cached_scores_only, steps, warmup, dev, world)
        ms_p = _timed_steps(plain, steps, warmup, dev, world)
        cands = sum(int(b[1].numel()) for b in bs) / 4
        r = imps(ms_c, B)
        r.update(news_table_encode={"news": M, "ms": enc_ms, "news_per_s": M / (enc_ms / 1e3)},
                 candidates_per_batch=cands, cached_scores_only=imps(ms_s, B), plain_eval_forward_scores_only=imps(ms_p, B),
                 note="value = cached scores + MRR/nDCG@5,10 (nrl_rank_metrics, one CTA per impression) + AUROC (torch.unique rank sums) per batch; "
                      "cached_scores_only / plain_eval_forward_scores_only compare the two ways of producing the scores",
                 workload=f"eval at MINDlarge-dev shape: B={B} whole impressions/GPU (ragged histories <= {HIST}, candidate lists "
                          f"lognormal up to 300), scores from news vectors cached once per epoch + AUC/MRR/nDCG on device; "
                          f"plain_eval_forward re-encodes every news of the batch")
        return r
    entry("eval_cached_mindlarge_dev", eval_cached)
    return out


def make_trainer(NRMSTrainer, params, dev, prec, world, mode):
    """Trainer + the name of the gradient exchange it uses.  One GPU: no exchange.  `auto`: the fused peer-memory
    kernel if every rank can map its peers AND two probe steps leave bit-identical replicas with no barrier
    timeout; otherwise NCCL all-reduce + dense Adam.  All ranks take the same branch (the verdict is all-reduced)."""
    import torch.distributed as dist
    if world == 1 or mode == "nccl":
        return NRMSTrainer(params, H, device=dev, dropout_p=DROPOUT, precision=prec, exchange="nccl"), \
            ("none (1 GPU)" if world == 1 else "nccl all-reduce + dense Adam")
    def all_ok(ok):
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())
    why, tr = "", None
    try:
        tr = NRMSTrainer(params, H, device=dev, dropout_p=DROPOUT, precision=prec, exchange="peer", exchange_timeout_s=5.0)
        ok = True
    except Exception as e:  # e.g. no peer access / IPC not permitted on this box
        ok, why = False, f"{type(e).__name__}: {e}"
    if all_ok(ok):
        # probe: two tiny steps, then the replicas must be the same bits on every rank and no barrier may have timed out
        rank = dist.get_rank()
        hb = make_batch(4, VOCAB, hist="fixed", max_hist=6, cand="train", seed=4321 + rank, max_title_len=L)
        b = {"x_hist": {"title": hb["x_hist"]["title"].to(dev)}, "x_cand": {"title": hb["x_cand"]["title"].to(dev)},
             "batch_hist": hb["batch_hist"].to(dev), "batch_cand": hb["batch_cand"].to(dev), "labels": hb["labels"].to(dev)}
        keep = tr.flat.clone()
        # probe step 1: a training step whose fused exchange is checked against torch.optim.Adam on the NCCL-averaged
        # gradients of the same backward pass (NRMSTrainer.probe_exchange)
        adam_ok, cleared = tr.probe_exchange(b, 4, 6, CAND)
        tr.train_step(b, 4, 6, CAND)  # probe step 2 through the public call
        status = tr.peer_block.status()
        lo, hi = tr.flat.clone(), tr.flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok = status == 0 and adam_ok and cleared and torch.equal(lo, hi) and not torch.equal(tr.flat, keep)
        if all_ok(ok):
            tr.flat.copy_(keep)  # back to the initial replica; the exchange epoch keeps counting
            tr.m.zero_()
            tr.v.zero_()
            tr.step_count = 0
            torch.cuda.synchronize()
            dist.barrier()
            return tr, ("peer (nrl_exchange_adam_step: row-sparse reduce-scatter + sharded Adam + all-gather + zero_grad in one "
                        "kernel over NVLink peer memory; probe: equals Adam on the NCCL-averaged gradients, replicas bit-identical)")
        why = (f"probe failed on some rank (here: status {status}, equals NCCL-mean Adam {adam_ok}, gradients cleared "
               f"{cleared})")
    if mode == "peer":
        raise SystemExit(f"--exchange peer: {why or 'a peer rank failed'}")
    dist.barrier()  # every rank, whether or not its own construction succeeded
    if tr is not None and tr.peer_block is not None:
        tr.peer_block.close()
    return NRMSTrainer(params, H, device=dev, dropout_p=DROPOUT, precision=prec, exchange="nccl"), \
        f"nccl all-reduce + dense Adam (peer exchange unavailable: {why or 'a peer rank failed'})"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default="nrms", choices=["nrms", "naml"],
                    help="nrms = the headline (BASELINE.json configs[1]); naml = configs[4] shape through NAMLModule")
    ap.add_argument("--exchange", default=os.environ.get("NRL_EXCHANGE", "auto"), choices=["auto", "nccl", "peer"],
                    help="N > 1: nccl = all-reduce + dense Adam; peer = nrl_exchange_adam_step over NVLink peer memory; "
                         "auto = peer if its self-check passes on this box, else nccl (the line says which)")
    ap.add_argument("--vocab", type=int, default=VOCAB, help="70000 = MINDsmall-shape (headline), 130000 = MINDlarge-shape")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (configs key)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    globals()["VOCAB"] = args.vocab
    globals()["WORKLOAD"] = WORKLOAD.replace("V=70000", f"V={args.vocab}")

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # rank 0's stdout carries ONE JSON line: NCCL writes its version banner to fd 1 whatever NCCL_DEBUG says,
        # so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    if args.model == "naml":
        return main_naml(args, rank, local_rank, world)

    import torch.distributed as dist
    from newsreclib_b200 import _lib, ops
    from newsreclib_b200.trainer import NRMSTrainer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: newsreclib_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout (rank 0 prints ONE JSON line)
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    prec = ops.PREC_BF16X3 if args.precision == "bf16x3" else ops.PREC_BF16

    B = args.batch
    params = make_nrms_params(VOCAB, E, H, Q, seed=1234)  # same init on every rank (DDP replica)
    trainer, exchange_used = make_trainer(NRMSTrainer, params, dev, prec, world, args.exchange)
    # a ring of distinct batches per rank so consecutive steps never see the same ids
    n_ring = 4
    host_batches, dev_batches = [], []
    for i in range(n_ring):
        hb = make_batch(B, VOCAB, hist="fixed", max_hist=HIST, cand="train", seed=1234 + rank * 100 + i, max_title_len=L)
        keep = {"x_hist": {"title": hb["x_hist"]["title"].pin_memory()}, "x_cand": {"title": hb["x_cand"]["title"].pin_memory()},
                "batch_hist": hb["batch_hist"].pin_memory(), "batch_cand": hb["batch_cand"].pin_memory(),
                "labels": hb["labels"].pin_memory()}
        host_batches.append(keep)
        dev_batches.append({"x_hist": {"title": keep["x_hist"]["title"].to(dev)}, "x_cand": {"title": keep["x_cand"]["title"].to(dev)},
                            "batch_hist": keep["batch_hist"].to(dev), "batch_cand": keep["batch_cand"].to(dev),
                            "labels": keep["labels"].to(dev)})
    Hmax, Cmax = HIST, CAND
    nh, nc = B * HIST, B * CAND

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    step_dev = lambda i: trainer.train_step(dev_batches[i % n_ring], B, Hmax, Cmax)
    scores_host = torch.empty(B, Cmax).pin_memory()
    loss_host = torch.empty(1).pin_memory()
    step_host = lambda i: trainer.train_step_host(host_batches[i % n_ring], B, Hmax, Cmax, scores_host, loss_host)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()  # started before the warm-up so nvidia-smi is already streaming
    for i in range(args.warmup):
        step_dev(i)
    torch.cuda.synchronize()

    # ---- timed region A: device-resident inputs ------------------------------------------
    l0 = lib.nrl_launch_count()
    t_host0 = time.perf_counter()
    ms_total = timed_region(step_dev, args.steps)
    t_host1 = time.perf_counter()
    launches = lib.nrl_launch_count() - l0
    if sampler:
        sampler.window(t_host0, t_host1)
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)

    # ---- timed region B: same steps with per-launch events -> roofline of the dominant kernel
    roof = None
    prof_steps = min(args.steps, 10)
    torch.cuda.synchronize()
    if rank == 0:
        lib.nrl_profile_start(torch.cuda.current_stream().cuda_stream)
    for i in range(prof_steps):  # every rank steps (the steps contain the gradient all-reduce); rank 0 records
        step_dev(i)
    torch.cuda.synchronize()
    if rank == 0:
        import ctypes as C
        maxrec, stride = 200 * prof_steps, 48
        names = C.create_string_buffer(maxrec * stride)
        msbuf = (C.c_float * maxrec)()
        n = lib.nrl_profile_stop(names, stride, msbuf, maxrec)
        agg = {}
        per_launch = {}  # name -> durations in launch order (all profiled steps)
        for i in range(n):
            nm = names.raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode()
            t, c = agg.get(nm, (0.0, 0))
            agg[nm] = (t + msbuf[i], c + 1)
            per_launch.setdefault(nm, []).append(msbuf[i])
        rows_news, rows_user = (nh + nc) * L, B * Hmax
        gemm_ms = sum(t for nm, (t, c) in agg.items() if nm.startswith("gemm")) / prof_steps
        gemm_launches = sum(c for nm, (t, c) in agg.items() if nm.startswith("gemm")) / prof_steps
        flops_step = sum(f() for f in GEMM_FLOPS.values()) * (rows_news + rows_user)
        hbm, tf_burst, tf_sust, src = peaks()
        achieved = flops_step / (gemm_ms / 1e3) / 1e12
        step_prof_ms = sum(t for t, c in agg.values()) / prof_steps
        top = sorted(((t / prof_steps, nm, c // prof_steps) for nm, (t, c) in agg.items()), reverse=True)[:14]
        roof = {"bound": "tensor", "kernel": "nrl_gemm_tc2_kernel / nrl_gemm_tc_kernel (all 18 tcgen05 GEMM launches of a step)",
                "achieved": achieved, "peak": tf_sust, "unit": "TFLOP/s", "frac": achieved / tf_sust,
                "peak_source": f"{src} bf16 dense, sustained (kernel timed inside a long step)",
                "traffic": gemm_traffic_per_launch(), "traffic_unit": "bytes of DRAM traffic per GEMM launch "
                "(ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the step's 18 launches; "
                "profiles/r02_gemm_traffic.json)",
                "algorithmic_gflop_per_step": flops_step / 1e9,
                "issued_passes": 3 if prec == ops.PREC_BF16X3 else 1,
                "kernel_ms_per_step": gemm_ms, "launches_per_step": gemm_launches,
                "share_of_step": gemm_ms / step_prof_ms,
                "top_kernels_ms_per_step": [[nm, round(t, 4), c] for t, nm, c in top]}
        # every GEMM of the title block on its own (the longer of a name's two launches per step): algorithmic FLOP/s
        # against the tensor peak (x issued passes = share of the pipe) AND algorithmic operand + result bytes against the
        # HBM peak -- the 300-wide reductions of the out-projection / additive GEMMs sit left of the ridge
        # (peak FLOP/s / peak B/s = 212 issued FLOP per byte): they are bound by their output bytes, not by the tensor pipe
        planes = 2 if prec == ops.PREC_BF16X3 else 1
        Ep_, Qp_, P3_, LDQ_ = (E + 1 + 15) // 16 * 16, (Q + 15) // 16 * 16, (3 * E + 15) // 16 * 16, (3 * E + 31) // 32 * 32
        mwb = 4 * ((E + 31) // 32)
        gemm_bytes = {"gemm in_proj": planes * 2 * Ep_ + 4 * LDQ_,
                      "gemm out_proj": planes * 2 * Ep_ + 4 * E + planes * 2 * Ep_ + mwb,
                      "gemm additive": planes * 2 * Ep_ + 4 * Q + 4,
                      "gemm additive dgrad": planes * 2 * Qp_ + 4 + mwb + planes * 2 * Ep_,
                      "gemm additive wgrad": planes * 2 * Qp_ + planes * 2 * Ep_,
                      "gemm out_proj dgrad": planes * 2 * Ep_ + 4 * E,
                      "gemm out_proj wgrad": planes * 2 * Ep_ + planes * 2 * Ep_,
                      "gemm in_proj dgrad": planes * 2 * P3_ + 4 * E + mwb,
                      "gemm in_proj wgrad": planes * 2 * P3_ + planes * 2 * Ep_}
        per_gemm = []
        for nm, f in GEMM_FLOPS.items():
            d = per_launch.get(nm, [])
            k = len(d) // prof_steps
            if k < 1:
                continue
            ms = sum(max(d[i * k:(i + 1) * k]) for i in range(prof_steps)) / prof_steps
            tf = rows_news * f() / (ms / 1e3) / 1e12
            gbs = rows_news * gemm_bytes[nm] / (ms / 1e3) / 1e9
            ft, fh = tf / tf_sust, gbs / hbm
            per_gemm.append({"kernel": nm + " (title block)", "ms": round(ms, 4), "algorithmic_tflops": round(tf, 1),
                             "frac_of_bf16_peak": round(ft, 3), "issued_frac": round(ft * roof["issued_passes"], 3),
                             "algorithmic_gbs": round(gbs, 1), "frac_of_hbm_peak": round(fh, 3),
                             "bound": "tensor" if ft * roof["issued_passes"] >= fh else "hbm"})
        roof["per_gemm"] = per_gemm

    # HBM-bound kernels: algorithmic bytes per step (per token row of the title block /
    # user block; Adam per parameter) over the live per-launch durations of region B
    hbm_kernels = None
    if rank == 0:
        n_params = sum(v.numel() for v in params.values())
        mw_bytes = 4 * ((E + 31) // 32)
        Qp = (Q + 15) // 16 * 16
        # bytes per token row of a MHSA + additive block (title block: (nh + nc) * L rows; user block: B * Hmax rows)
        per_row = {"attn_fwd": 3 * 4 * E + 2 * 2 * (E + 4) + 4 * H,
                   "attn_bwd": 3 * 4 * E + 4 * E + 4 * H + 2 * 2 * (3 * E + 12),
                   "pool_fwd": 4 * E + 8, "pool_bwd": 4 * E + 4 * Q + 4 + 2 * 2 * Qp}
        per_group = {"pool_fwd": 4 * E, "pool_bwd": 4 * E}
        hbm_kernels = []

        def add(label, ms, nbytes, **extra):
            gbs = nbytes / (ms / 1e3) / 1e9
            hbm_kernels.append({"kernel": label, "ms_per_step": round(ms, 4), "algorithmic_mbytes": round(nbytes / 1e6, 1),
                                "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm, 3), **extra})
        if "gather_split" in agg:
            add("gather_split", agg["gather_split"][0] / prof_steps, rows_news * (8 + 4 * E + mw_bytes + 2 * 2 * (E + 4)))
        for nm in ("attn_fwd", "attn_bwd", "pool_fwd", "pool_bwd"):
            # two launches per step: the title block ((nh + nc) * L = 105 600 rows) and the user block (B * Hmax = 3 200 rows:
            # 33x less data, a latency-bound launch).  Launch order differs between forward and backward: the longer one of a
            # step is the title block.
            d = per_launch.get(nm, [])
            k = len(d) // prof_steps
            if k != 2:
                continue
            big = sum(max(d[i * k:(i + 1) * k]) for i in range(prof_steps)) / prof_steps
            small = sum(min(d[i * k:(i + 1) * k]) for i in range(prof_steps)) / prof_steps
            add(nm + " (title block)", big, rows_news * per_row[nm] + (nh + nc) * per_group.get(nm, 0))
            add(nm + " (user block, 3 200 rows: latency-bound)", small, rows_user * per_row[nm] + B * per_group.get(nm, 0))
        if "emb_grad" in agg:
            add("emb_grad", agg["emb_grad"][0] / prof_steps, rows_news * (4 * E + 8 + 2 * 4 * E),
                note="the read-modify-write of the table rows is served by L2 (ncu: DRAM 15 %): this is an L2-assisted figure")
        if "adam" in agg:
            add("adam", agg["adam"][0] / prof_steps, n_params * 28)

    # the fused exchange kernel (N > 1, --exchange peer): bytes that cross NVLink per rank and step, per direction
    # (gradients of the owned slice pulled from W-1 peers; new parameters of the owned slice pushed to W-1 peers;
    # a rank also serves the same amounts to its peers), over the live launch duration -- which includes waiting
    # for the slowest rank's backward pass at the ready barrier
    exchange_info = None
    try:
        if rank == 0 and world > 1 and "exchange_adam" in agg:
            n_flat = trainer.flat.numel()
            ms_x = agg["exchange_adam"][0] / prof_steps
            dense = 4.0 * n_flat * (world - 1) / world      # new parameters of the owned slice to W-1 peers (and received)
            # gradients: only the table rows a rank touched cross the links (plus the dense non-embedding parameters)
            touched = sum(int(torch.unique(torch.cat([hb["x_hist"]["title"].reshape(-1), hb["x_cand"]["title"].reshape(-1)])).numel())
                          for hb in host_batches) / len(host_batches)
            n_table = trainer.table.numel()
            grads = 4.0 * ((n_flat - n_table) + touched * E) * (world - 1) / world
            exchange_info = {"kernel": "exchange_adam", "ms_per_step": round(ms_x, 4),
                             "nvlink_mbytes_params_per_direction": round(dense / 1e6, 1),
                             "nvlink_mbytes_grads_per_direction": round(grads / 1e6, 1),
                             "nvlink_mbytes_grads_if_dense": round(dense / 1e6, 1),
                             "table_rows_touched_per_rank": round(touched), "table_rows": int(trainer.table.shape[0]),
                             "achieved_gbs_per_direction": round((dense + grads) / (ms_x / 1e3) / 1e9, 1),
                             "note": "kernel time includes the local non-zero-row scan (84 MB of HBM reads), the wait for the "
                                     "slowest rank at the ready barrier and the sharded Adam",
                             "adam_elements_per_rank": n_flat // world,
                             "timeline_us_last_step_rank0": trainer.peer_block.timeline_us()}
    except Exception as e:  # never lose the line over a diagnostic
        exchange_info = {"error": repr(e)}

    # ---- timed region C: end to end through the C ABI with host buffers -------------------
    for i in range(2):
        step_host(i)
    ms_e2e = timed_region(step_host, args.steps) / args.steps
    e2e_val = world * B / (ms_e2e / 1e3)
    h2d = (nh + nc) * L * 8 + (nh + nc) * 8 + nc * 4
    d2h = B * Cmax * 4 + 4

    # ---- timed region D: eval forward only (scores + loss, no backward / Adam), device-resident inputs
    for i in range(2):
        trainer.eval_forward(dev_batches[i % n_ring], B, Hmax, Cmax)
    ms_eval = timed_region(lambda i: trainer.eval_forward(dev_batches[i % n_ring], B, Hmax, Cmax), args.steps) / args.steps

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        bs, timed, threads, kind = cpu_train_steps(B, 6, 1, budget_s=25.0)
        cms = 1e3 * sum(timed) / len(timed)
        cpu = {"value": bs / (cms / 1e3), "unit": "impressions/s", "cores": threads, "kind": kind,
               "sample": f"{len(timed)} train steps of {bs} impressions on the host CPU, fp32, {threads} threads: "
                         f"{CPU_KIND_TEXT[kind]}",
               "ms_per_step": cms}

    if trainer.peer_block is not None and trainer.peer_block.status() != 0:
        raise SystemExit(f"rank {rank}: a peer-exchange barrier timed out (code {trainer.peer_block.status()}): numbers invalid")
    extras = None
    if not args.no_extras:
        extras = extra_configs(args, dev, rank, world, exchange_used)
    if rank == 0:
        out = {
            "metric": "impressions/sec", "value": value, "unit": "impressions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32 via bf16x3 split on tcgen05 (fp32 accumulate)" if prec == ops.PREC_BF16X3 else "bf16 (fp32 accumulate)",
            "data": "synthetic",
            "config": run_config(world, B),
            "impl_detail": {"exchange": exchange_used, "precision": args.precision},
            "e2e": {"value": e2e_val, "unit": "impressions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e},
            "eval_forward": {"value": world * B / (ms_eval / 1e3), "unit": "impressions/s", "ms_per_step": ms_eval},
            "gpu_launches": int(launches),
            "clocks": clocks, "roofline": roof, "hbm_kernels": hbm_kernels, "exchange": exchange_info,
            "cpu_baseline": cpu,
            "configs": extras,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
