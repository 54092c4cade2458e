"""Helpers of the transformer tests: fixture loading, the CUDA path through the C ABI, the
oracle with autograd.  Test infrastructure (may import the oracle)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

EMBED_KEYS = ("embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight",
              "embeddings.token_type_embeddings.weight", "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias")
LAYER_KEYS = ("attention.self.query.weight", "attention.self.query.bias", "attention.self.key.weight",
              "attention.self.key.bias", "attention.self.value.weight", "attention.self.value.bias",
              "attention.output.dense.weight", "attention.output.dense.bias", "attention.output.LayerNorm.weight",
              "attention.output.LayerNorm.bias", "intermediate.dense.weight", "intermediate.dense.bias",
              "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias")


def load_tfm_golden(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    hidden, heads, inter, layers, vocab, N, T, max_pos = [int(x) for x in g["meta"]]
    cfg = dict(hidden=hidden, heads=heads, inter=inter, layers=layers, vocab=vocab, N=N, T=T, max_pos=max_pos,
               eps=float(g["eps"]), frozen=[int(x) for x in g["frozen"]], bert=bool(int(g["bert"])) if "bert" in g else False)
    params = {k[len("param/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param/")}
    grads = {k[len("grad/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("grad/")}
    return g, cfg, params, grads


def random_tfm_params(hidden, heads, inter, layers, vocab, max_pos, seed, wstd=0.06):
    """HF-named RoBERTa parameters with O(1) attention logits and GELU inputs (HF's own init, std 0.02 and zero biases,
    leaves most terms barely exercised)."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std):
        return torch.randn(*shape, generator=g) * std
    P = {"embeddings.word_embeddings.weight": rn(vocab, hidden, std=0.5),
         "embeddings.position_embeddings.weight": rn(max_pos, hidden, std=0.2),
         "embeddings.token_type_embeddings.weight": rn(1, hidden, std=0.2),
         "embeddings.LayerNorm.weight": 1 + rn(hidden, std=0.1), "embeddings.LayerNorm.bias": rn(hidden, std=0.1)}
    for l in range(layers):
        pre = f"encoder.layer.{l}."
        for nm, (o, i) in {"attention.self.query": (hidden, hidden), "attention.self.key": (hidden, hidden),
                           "attention.self.value": (hidden, hidden), "attention.output.dense": (hidden, hidden),
                           "intermediate.dense": (inter, hidden), "output.dense": (hidden, inter)}.items():
            P[pre + nm + ".weight"] = rn(o, i, std=wstd)
            P[pre + nm + ".bias"] = rn(o, std=0.1)
        for nm in ("attention.output.LayerNorm", "output.LayerNorm"):
            P[pre + nm + ".weight"] = 1 + rn(hidden, std=0.1)
            P[pre + nm + ".bias"] = rn(hidden, std=0.1)
    return P


def random_text(N, T, vocab, seed, min_len=3):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min_len, T + 1, (N,), generator=g)
    lens[0] = T
    ids = torch.randint(3, vocab, (N, T), generator=g)
    att = (torch.arange(T)[None, :] < lens[:, None]).long()
    ids = torch.where(att.bool(), ids, torch.ones_like(ids))
    ids[:, 0] = 0
    return ids, att


def param_list(P, layers, dev="cuda", frozen=(), train_embed=True):
    """The tensors ops.TfmEncoderFn takes (leaves on `dev`), with the frozen layers' requires_grad off."""
    out, names = [], []
    for k in EMBED_KEYS:
        t = P[k][0] if k.endswith("token_type_embeddings.weight") else P[k]
        out.append(t.detach().clone().to(dev).requires_grad_(train_embed))
        names.append(k)
    for l in range(layers):
        for k in LAYER_KEYS:
            out.append(P[f"encoder.layer.{l}.{k}"].detach().clone().to(dev).requires_grad_(l not in frozen))
            names.append(f"encoder.layer.{l}.{k}")
    return out, names


def gpu_tfm(P, cfg, ids, att, w=None, frozen=(), train_embed=True, training=False, seed=0, p_hidden=0.0, p_attn=0.0,
            precision=None):
    """The CUDA path through ops.TfmEncoderFn -> the C ABI; returns out, {name: grad} (grads if w is given)."""
    from newsreclib_b200 import ops
    bert = bool(cfg.get("bert", False))
    st = ops.TfmState(cfg["hidden"], cfg["heads"], cfg["inter"], cfg["layers"], cfg["vocab"], cfg["max_pos"], 0 if bert else 1,
                      cfg["eps"], p_hidden, p_attn, position_mode=1 if bert else 0)
    leaves, names = param_list(P, cfg["layers"], "cuda", frozen, train_embed)
    out = ops.TfmEncoderFn.apply(ids.cuda(), None if att is None else att.cuda(), st, training, seed,
                                 ops.PREC_BF16X3 if precision is None else precision, *leaves)
    grads = {}
    if w is not None:
        (out * w.cuda()).sum().backward()
        for n, t in zip(names, leaves):
            if t.grad is not None:
                grads[n] = t.grad.detach().cpu()
    torch.cuda.synchronize()
    ops.device_status()
    return out.detach().cpu(), grads


def oracle_tfm(P, cfg, ids, att, w=None, frozen=(), train_embed=True, masks=None, p_hidden=0.0, p_attn=0.0,
               dtype=torch.float32):
    from oracle import tfm_oracle as TO
    Q = {}
    for k, v in P.items():
        layer = int(k.split(".")[2]) if k.startswith("encoder.layer.") else -1
        rg = (layer >= 0 and layer not in frozen) or (layer < 0 and train_embed)
        Q[k] = v.detach().clone().to(dtype).requires_grad_(rg and w is not None)
    bert = bool(cfg.get("bert", False))
    out = TO.encoder(ids, att, Q, cfg["heads"], cfg["layers"], pad_idx=0 if bert else 1, eps=cfg["eps"], masks=masks,
                     p_hidden=p_hidden, p_attn=p_attn, bert=bert)
    grads = {}
    if w is not None:
        (out * w.to(dtype)).sum().backward()
        for k, v in Q.items():
            if v.grad is not None:
                grads[k] = v.grad[0].detach() if k.endswith("token_type_embeddings.weight") else v.grad.detach()
    return out.detach(), grads


def grad_errors(got, ref):
    """max |got - ref| per tensor relative to max |ref|.  Two guards against ill-posed denominators: a floor of 1e-4 of
    the largest reference gradient, and the KEY-BIAS gradient -- exactly zero in exact arithmetic (adding a constant to
    every key's logit of a query leaves the softmax unchanged), i.e. a sum over all tokens of dK rows that cancels to
    rounding noise in any finite precision -- is measured against the query-bias gradient of the same layer (the same sum
    without the cancellation)."""
    floor = 1e-4 * max(float(v.abs().max()) for v in ref.values())
    errs = {}
    for k, r in ref.items():
        r2 = r[0] if (k.endswith("token_type_embeddings.weight") and r.dim() == 2) else r
        assert k in got, f"no gradient for {k}"
        den = max(float(r2.abs().max()), floor)
        if k.endswith("attention.self.key.bias"):
            den = max(den, float(ref[k.replace(".key.bias", ".query.bias")].abs().max()))
        errs[k] = float((got[k].double() - r2.double()).abs().max() / den)
    return errs
