"""Pin ``oracle/tfm_oracle.py`` against the real thing and mint the transformer fixtures.

Runs ONLY in the build container: it imports the unmodified reference ``PLM`` class from ``/root/reference``
(``newsreclib/models/components/encoders/news/text.py:15-109``), lets ITS constructor load a random-init RoBERTa
saved offline (``AutoModel.from_pretrained``, ``:67``) and freeze layers by name (``:70-73``), and runs
``ref.plm_model(**text)[0]`` (``:92``) -- HuggingFace ``transformers`` (third-party; version recorded in the fixture)
-- forward and backward.  The same inputs go through the restatement; agreement is asserted and inputs, weights,
outputs and every gradient are written to ``tests/golden/tfm_*.npz``.

Usage:  python oracle/make_tfm_golden.py            (mint)
        python oracle/make_tfm_golden.py --check    (re-run and compare with the committed fixtures)
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import tfm_oracle as TO  # noqa: E402
from oracle._golden_io import save as golden_save  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make_text(N, T, vocab, seed, full_rows=1, pad=1, cls=0):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(3, T + 1, (N,), generator=g)
    lens[:full_rows] = T
    ids = torch.randint(3, vocab, (N, T), generator=g)
    att = (torch.arange(T)[None, :] < lens[:, None]).long()
    ids = torch.where(att.bool(), ids, torch.full_like(ids, pad))  # pad_token_id
    ids[:, 0] = cls  # <s> / [CLS]
    return {"input_ids": ids, "attention_mask": att}, g


def tfm_case(name, hidden, heads, inter, layers, vocab, N, T, frozen, seed, bert=False):
    import transformers
    from transformers import BertConfig, BertModel, RobertaConfig, RobertaModel
    from newsreclib.models.components.encoders.news.text import PLM
    torch.manual_seed(seed)
    pad = 0 if bert else 1
    kw = dict(vocab_size=vocab, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
              intermediate_size=inter, max_position_embeddings=T + 4, pad_token_id=pad)
    cfg = BertConfig(type_vocab_size=2, **kw) if bert else RobertaConfig(type_vocab_size=1, **kw)
    with tempfile.TemporaryDirectory() as d:
        m0 = BertModel(cfg) if bert else RobertaModel(cfg)
        with torch.no_grad():  # HF initialises biases and LayerNorm to 0 / 1: perturb so that every term is exercised
            for n, p in m0.named_parameters():
                if p.dim() == 1:
                    p.add_(0.05 * torch.randn_like(p))
                else:
                    p.mul_(3.0)  # std 0.06: attention logits and GELU inputs of O(1)
        m0.save_pretrained(d)
        ref = PLM(plm_model=d, frozen_layers=frozen, embed_dim=hidden, use_mhsa=True, apply_reduce_dim=False,
                  reduced_embed_dim=None, num_heads=2, query_dim=8, dropout_probability=0.2).eval()
    tf = ref.plm_model
    text, g = make_text(N, T, vocab, seed, pad=pad, cls=2 if bert else 0)
    out = tf(**text)[0]
    w = torch.randn(N, T, hidden, generator=g)
    (out * w).sum().backward()
    sd = {k: v.detach().clone() for k, v in tf.state_dict().items() if not k.startswith("pooler.")}
    rgrads = {k: p.grad.detach().clone() for k, p in tf.named_parameters() if p.grad is not None}
    frozen_names = [k for k, p in tf.named_parameters() if not p.requires_grad]
    assert all(any(f"layer.{l}." in k for l in frozen) for k in frozen_names) and (not frozen or frozen_names)
    # the restatement on the same inputs
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    oo = TO.encoder(text["input_ids"], text["attention_mask"], P, heads, layers, pad_idx=pad, eps=cfg.layer_norm_eps, bert=bert)
    (oo * w).sum().backward()
    worst = rel(oo.detach(), out.detach())
    assert worst < 2e-5, worst
    # the key-bias gradient is exactly zero in exact arithmetic (a per-query shift of the logits leaves the softmax
    # unchanged); in fp32 it is rounding noise, so errors are measured against a floor of 1e-4 of the largest gradient
    floor = 1e-4 * max(float(v.abs().max()) for v in rgrads.values())
    for k, gref in rgrads.items():
        if P[k].grad is None:  # BERT: rows 1.. of the token-type table are never used (token_type_ids are all 0)
            assert float(gref.abs().max()) == 0.0, k
            continue
        r = float((P[k].grad - gref).abs().max() / max(float(gref.abs().max()), floor))
        worst = max(worst, r)
        assert r < 5e-4, (k, r)
    rec = dict(meta=np.array([hidden, heads, inter, layers, vocab, N, T, T + 4]), bert=np.array(int(bert)),
               eps=np.array(cfg.layer_norm_eps), frozen=np.array(frozen, dtype=np.int64),
               input_ids=text["input_ids"].numpy(), attention_mask=text["attention_mask"].numpy(), w=w.numpy(),
               out=out.detach().numpy(), oracle_vs_reference_maxrel=np.array(worst),
               transformers_version=np.array(transformers.__version__))
    rec.update({"param/" + k: v.numpy() for k, v in sd.items() if v.is_floating_point()})
    rec.update({"grad/" + k: v.numpy() for k, v in rgrads.items()})
    golden_save(os.path.join(GOLD, name + ".npz"), **rec)
    print(f"{name}: restatement vs HF {type(tf).__name__} (through the reference PLM class) max rel {worst:.2e}; "
          f"{len(rgrads)} gradients, {len(frozen_names)} frozen tensors")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(4)
    # two layers, layer 0 frozen (the reference's name match), embeddings trainable; ragged titles with pads
    tfm_case("tfm_tiny", hidden=128, heads=2, inter=256, layers=2, vocab=120, N=7, T=12, frozen=[0], seed=11)
    # T > 32 (two key blocks), three layers with the middle one frozen only by accident of the list
    tfm_case("tfm_t40", hidden=128, heads=2, inter=192, layers=3, vocab=90, N=5, T=40, frozen=[0, 1], seed=12)
    # BertModel (bert-base style PLMs): positions 0..T-1, padding row 0 of the word table only, two token-type rows
    tfm_case("tfm_bert", hidden=128, heads=2, inter=256, layers=2, vocab=100, N=6, T=14, frozen=[0], seed=13, bert=True)
