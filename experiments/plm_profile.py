"""Per-launch timing of one NRMS-PLM training step (roberta-base shape) on the sm_100a path: which of this library's
kernels the step spends its time in (nrl_profile_start / stop: one CUDA event per launch on the launching stream) and how
much of the wall time of the step is outside them (torch glue, optimizer).  Usage: python experiments/plm_profile.py [T]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import _dev_batch, _module_kwargs, HIST, CAND, L, Q  # noqa: E402
from newsreclib_b200 import _lib  # noqa: E402
from newsreclib_b200.synthetic import make_batch  # noqa: E402
from newsreclib_b200.trainer import ModuleTrainer  # noqa: E402


def main():
    from transformers import RobertaConfig, RobertaModel
    from newsreclib_b200.models.general_rec.nrms_module import NRMSModule
    max_len = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    dev = torch.device("cuda:0")
    lib = _lib.load()
    outputs = {"train": ["preds", "targets", "cand_news_size"], "val": ["preds", "targets", "cand_news_size"],
               "test": ["preds", "targets", "cand_news_size"]}
    torch.manual_seed(1234)
    plm = RobertaModel(RobertaConfig(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1), add_pooling_layer=False)
    m = NRMSModule(dataset_attributes=["title", "category"], attributes2encode=["title"], use_plm=True,
                   pretrained_embeddings_path=None, plm_model=plm, frozen_layers=list(range(8)), embed_dim=768, num_heads=16,
                   query_dim=Q, **_module_kwargs(outputs))
    tr = ModuleTrainer(m.to(dev), lr=1e-5, exchange="nccl")
    rng = np.random.default_rng(1234)

    def plm_news(n):
        lens = np.clip(rng.poisson(16, n), 6, max_len)
        if max_len == 96:
            lens[0] = 96
        T = int(lens.max())
        ids = rng.integers(3, 50265, (n, T))
        mask = np.arange(T)[None, :] < lens[:, None]
        ids[~mask] = 1
        return {"input_ids": torch.from_numpy(ids).to(dev), "attention_mask": torch.from_numpy(mask.astype(np.int64)).to(dev)}
    b = _dev_batch(make_batch(8, 1000, hist="fixed", max_hist=HIST, cand="train", seed=40, max_title_len=L), dev)
    b["x_hist"]["title"], b["x_cand"]["title"] = plm_news(8 * HIST), plm_news(8 * CAND)
    for _ in range(3):
        tr.train_step(b)
    torch.cuda.synchronize()
    os.environ["NRL_WGRAD_STREAM"] = "0"
    steps = 3
    lib.nrl_profile_start(torch.cuda.current_stream().cuda_stream)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.train_step(b)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps * 1e3
    maxrec, stride = 4000 * steps, 48
    names = C.create_string_buffer(maxrec * stride)
    msbuf = (C.c_float * maxrec)()
    n = lib.nrl_profile_stop(names, stride, msbuf, maxrec)
    agg = {}
    for i in range(n):
        nm = names.raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode()
        t, c = agg.get(nm, (0.0, 0))
        agg[nm] = (t + msbuf[i], c + 1)
    tot = sum(t for t, c in agg.values()) / steps
    print(f"T = {b['x_hist']['title']['input_ids'].shape[1]} / {b['x_cand']['title']['input_ids'].shape[1]}; wall {wall:.2f} ms/step; "
          f"sum of this library's launches (incl. gaps before them) {tot:.2f} ms/step; {n // steps} launches/step")
    for nm, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"  {nm:36s} {t / steps:8.3f} ms  x{c // steps}")


if __name__ == "__main__":
    main()
