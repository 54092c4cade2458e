"""Per-launch timing of the drop-in NRMSModule path (model_step + autograd + ModuleTrainer) beside its wall time:
where the gap to the fused nrl_nrms_step comes from.  Usage: python experiments/module_profile.py"""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import _dev_batch, _module_kwargs, HIST, L, Q, E, H, VOCAB  # noqa: E402
from newsreclib_b200 import _lib  # noqa: E402
from newsreclib_b200.synthetic import make_batch, make_nrms_params  # noqa: E402
from newsreclib_b200.trainer import ModuleTrainer  # noqa: E402


def main():
    from newsreclib_b200.models.general_rec.nrms_module import NRMSModule
    dev = torch.device("cuda:0")
    lib = _lib.load()
    outputs = {"train": ["preds", "targets", "cand_news_size"], "val": ["preds", "targets", "cand_news_size"],
               "test": ["preds", "targets", "cand_news_size"]}
    naml = len(sys.argv) > 1 and sys.argv[1] == "naml"
    if naml:
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
        bs = [_dev_batch(make_batch(64, V, hist="fixed", max_hist=HIST, cand="train", seed=700 + i, max_title_len=L,
                                    abstract_len=LA), dev) for i in range(4)]
    else:
        params = make_nrms_params(VOCAB, E, H, Q, seed=1234)
        m = NRMSModule(dataset_attributes=["title", "category"], attributes2encode=["title"], use_plm=False,
                       pretrained_embeddings_path=None, plm_model=None, frozen_layers=None, embed_dim=E, num_heads=H,
                       query_dim=Q, pretrained_embeddings=params["news_encoder.text_encoders.title.embedding_layer.weight"],
                       **_module_kwargs(outputs))
        m.load_state_dict({k: v for k, v in params.items() if k in m.state_dict()})
        bs = [_dev_batch(make_batch(64, VOCAB, hist="fixed", max_hist=HIST, cand="train", seed=900 + i, max_title_len=L), dev)
              for i in range(4)]
    tr = ModuleTrainer(m.to(dev).train(), lr=1e-4, exchange="nccl")
    for i in range(5):
        tr.train_step(bs[i % 4])
    torch.cuda.synchronize()
    steps = 10
    t0 = time.perf_counter()
    for i in range(steps):
        tr.train_step(bs[i % 4])
    torch.cuda.synchronize()
    wall_free = (time.perf_counter() - t0) / steps * 1e3
    os.environ["NRL_WGRAD_STREAM"] = "0"
    lib.nrl_profile_start(torch.cuda.current_stream().cuda_stream)
    t0 = time.perf_counter()
    for i in range(steps):
        tr.train_step(bs[i % 4])
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps * 1e3
    maxrec, stride = 400 * steps, 48
    names = C.create_string_buffer(maxrec * stride)
    msbuf = (C.c_float * maxrec)()
    n = lib.nrl_profile_stop(names, stride, msbuf, maxrec)
    agg = {}
    for i in range(n):
        nm = names.raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode()
        t, c = agg.get(nm, (0.0, 0))
        agg[nm] = (t + msbuf[i], c + 1)
    tot = sum(t for t, c in agg.values()) / steps
    print(f"wall {wall_free:.3f} ms/step unprofiled, {wall:.3f} profiled; this library's launches (incl. the gaps before them) "
          f"{tot:.3f} ms/step; {n // steps} launches/step")
    for nm, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
        print(f"  {nm:36s} {t / steps:8.4f} ms  x{c // steps}")


if __name__ == "__main__":
    main()
