"""Seeded synthetic MIND-shaped ``RecommendationBatch`` generator.

Produces exactly the layout ``DatasetCollate.__call__`` builds
(reference ``newsreclib/data/components/rec_dataset.py:148-168,289-293``): ragged
PyG-style batches with sorted segment ids, titles right-padded with id 0 to
``max_title_len`` and float labels.  CPU tensors; callers move them to the device.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch


def _zipf_ids(rng: np.random.Generator, n: int, vocab: int, s: float = 1.07) -> np.ndarray:
    """ids ~ Zipf(s) over [1, vocab] by inverse-CDF on the truncated distribution."""
    ranks = np.arange(1, vocab + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** (-s))
    cdf /= cdf[-1]
    return (np.searchsorted(cdf, rng.random(n)) + 1).astype(np.int64)


def make_titles(rng: np.random.Generator, n_news: int, vocab: int, max_len: int = 30,
                mean_len: float = 11.5, min_len: int = 3) -> np.ndarray:
    lens = np.clip(rng.poisson(mean_len, n_news), min_len, max_len)
    ids = _zipf_ids(rng, n_news * max_len, vocab).reshape(n_news, max_len)
    ids[np.arange(max_len)[None, :] >= lens[:, None]] = 0
    return ids


def make_batch(
    batch_size: int = 64,
    vocab: int = 70000,
    hist: str = "fixed",
    max_hist: int = 50,
    cand: str = "train",
    max_title_len: int = 30,
    seed: int = 1234,
    abstract_len: Optional[int] = None,
    num_categories: int = 18,
) -> Dict:
    """hist: "fixed" (= max_hist) | "ragged" (clip(round(lognormal(3.2, .8)), 1, max_hist)).
    cand: "train" (1 positive + 4 negatives) | "eval" (clip(round(lognormal(3.3, .7)), 2, 300))."""
    rng = np.random.default_rng(seed)
    if hist == "fixed":
        h = np.full(batch_size, max_hist, dtype=np.int64)
    else:
        h = np.clip(np.round(rng.lognormal(3.2, 0.8, batch_size)), 1, max_hist).astype(np.int64)
    if cand == "train":
        c = np.full(batch_size, 5, dtype=np.int64)
    else:
        c = np.clip(np.round(rng.lognormal(3.3, 0.7, batch_size)), 2, 300).astype(np.int64)
    n_h, n_c = int(h.sum()), int(c.sum())

    labels = np.zeros(n_c, dtype=np.float32)
    off = np.concatenate([[0], np.cumsum(c)])
    for b in range(batch_size):
        labels[off[b] + rng.integers(0, c[b])] = 1.0

    def news(n: int) -> Dict:
        d = {
            "news_ids": torch.from_numpy(rng.integers(1, 10**6, n).astype(np.int64)),
            "title": torch.from_numpy(make_titles(rng, n, vocab, max_title_len)),
            "category": torch.from_numpy(rng.integers(1, num_categories + 1, n).astype(np.int64)),
            "subcategory": torch.from_numpy(rng.integers(1, 200, n).astype(np.int64)),
            "sentiment": torch.from_numpy(rng.integers(1, 4, n).astype(np.int64)),
            "sentiment_score": torch.from_numpy(rng.random(n).astype(np.float32)),
        }
        if abstract_len is not None:
            d["abstract"] = torch.from_numpy(
                make_titles(rng, n, vocab, abstract_len, mean_len=40.0, min_len=0))
        return d

    return {
        "batch_hist": torch.from_numpy(np.repeat(np.arange(batch_size), h)),
        "batch_cand": torch.from_numpy(np.repeat(np.arange(batch_size), c)),
        "x_hist": news(n_h),
        "x_cand": news(n_c),
        "labels": torch.from_numpy(labels),
        "user_ids": torch.from_numpy(rng.integers(1, 10**6, batch_size).astype(np.int64)),
        "user_idx": torch.arange(batch_size, dtype=torch.int64),
        # host-known dense widths (Hmax, Cmax): lets the drop-in modules skip the device sync of to_dense_batch
        "dense_widths": (int(h.max()), int(c.max())),
    }


def make_nrms_params(vocab: int, embed_dim: int = 300, num_heads: int = 15,
                     query_dim: int = 200, seed: int = 1234,
                     scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Random-init NRMS parameters under the reference ``state_dict`` key names
    (SURVEY.md §8b).  Table ~ N(0,1) (the reference's init for nearly every row,
    ``data_utils.py:56``); projections ~ U(-a, a) with a = 1/sqrt(fan_in) (``nn.Linear``
    default scale); query ~ U(-0.1, 0.1) (``layers/attention.py:22``)."""
    g = torch.Generator().manual_seed(seed)
    E, Q = embed_dim, query_dim

    def u(*shape, a):
        return (torch.rand(*shape, generator=g) * 2 - 1) * a * scale

    def block(prefix):
        a = 1.0 / (E ** 0.5)
        return {
            prefix + "multihead_attention.in_proj_weight": u(3 * E, E, a=(6.0 / (4 * E)) ** 0.5),
            prefix + "multihead_attention.in_proj_bias": u(3 * E, a=0.05),
            prefix + "multihead_attention.out_proj.weight": u(E, E, a=a),
            prefix + "multihead_attention.out_proj.bias": u(E, a=0.05),
            prefix + "additive_attention.linear.weight": u(Q, E, a=a),
            prefix + "additive_attention.linear.bias": u(Q, a=a),
            prefix + "additive_attention.query": u(Q, a=0.1),
        }

    p = {"news_encoder.text_encoders.title.embedding_layer.weight":
         torch.randn(vocab + 1, E, generator=g)}
    p.update(block("news_encoder.text_encoders.title."))
    p.update(block("user_encoder."))
    return p


def make_naml_params(vocab: int, embed_dim: int = 300, num_filters: int = 400, window: int = 3,
                     query_dim: int = 200, categ_embed_dim: int = 100, num_categories: int = 19,
                     seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Random-init NAML parameters under the reference ``state_dict`` key names (SURVEY.md §8b;
    the ``abstract`` prefix aliases the ``title`` text encoder and is not repeated).  Table ~ N(0,1);
    conv / linear ~ U(-a, a), a = 1/sqrt(fan_in); queries ~ U(-0.1, 0.1); category table ~ N(0,1)
    with row 0 zero (``nn.Embedding(padding_idx=0)``, ``category.py:56-58``)."""
    g = torch.Generator().manual_seed(seed)
    E, F_, Q, CE = embed_dim, num_filters, query_dim, categ_embed_dim

    def u(*shape, a):
        return (torch.rand(*shape, generator=g) * 2 - 1) * a

    def add(prefix, dim):
        a = 1.0 / (dim ** 0.5)
        return {prefix + "linear.weight": u(Q, dim, a=a), prefix + "linear.bias": u(Q, a=a),
                prefix + "query": u(Q, a=0.1)}

    t = "news_encoder.text_encoders.title."
    c = "news_encoder.category_encoders.category."
    a_conv = 1.0 / ((window * E) ** 0.5)
    p = {t + "embedding_layer.weight": torch.randn(vocab + 1, E, generator=g),
         t + "cnn.weight": u(F_, 1, window, E, a=a_conv), t + "cnn.bias": u(F_, a=a_conv)}
    p.update(add(t + "additive_attention.", F_))
    ctab = torch.randn(num_categories, CE, generator=g)
    ctab[0] = 0
    p[c + "embedding_layer.weight"] = ctab
    p[c + "linear.weight"] = u(F_, CE, a=1.0 / (CE ** 0.5))
    p[c + "linear.bias"] = u(F_, a=1.0 / (CE ** 0.5))
    p.update(add("news_encoder.combine_layer.", F_))
    p.update(add("user_encoder.additive_attention.", F_))
    return p
