"""Pin ``oracle/collate_oracle.py`` against the reference's OWN collate and mint ``tests/golden/collate_ref.npz``.

Runs ONLY in the build container (imports the unmodified reference from ``/root/reference``).
``newsreclib/data/components/rec_dataset.py`` imports ``mind_dataframe.py`` for the base class of its two
``Dataset``s, and that module needs ``omegaconf`` / ``hydra`` / the download utilities, none of which the collate
uses.  The import is satisfied with a stand-in module whose ``MINDDataFrame`` is ``torch.utils.data.Dataset``;
``rec_dataset.py`` itself is loaded unmodified, so ``RecommendationDatasetTest.__getitem__`` (``:102-116``),
``RecommendationDatasetTrain.__getitem__`` / ``_sample_candidates`` (``:39-95``, under a fixed ``np.random`` seed)
and ``DatasetCollate.__call__`` (``:148-293``) below are the reference's own code.

The fixture stores the synthetic news table and behaviours (plain arrays), the per-impression row lists the
datasets produced, and every tensor of the resulting ``RecommendationBatch``es (test-split and train-split with
4:1 negative sampling).  Usage:  python oracle/make_collate_golden.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

stub = types.ModuleType("newsreclib.data.components.mind_dataframe")
stub.MINDDataFrame = torch.utils.data.Dataset
sys.modules["newsreclib.data.components.mind_dataframe"] = stub
from newsreclib.data.components.rec_dataset import (  # noqa: E402
    DatasetCollate, RecommendationDatasetTest, RecommendationDatasetTrain)

from oracle import collate_oracle as CO  # noqa: E402
from oracle._golden_io import save as golden_save  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
L_TITLE, L_ABS = 30, 50


def make_tables(seed=2024, n_news=60, n_imp=9):
    rng = np.random.default_rng(seed)
    nid_num = rng.permutation(np.arange(1000, 1000 + 5 * n_news))[:n_news]
    news = {
        "nid": nid_num.tolist(),
        # lengths 0 .. 45: empty titles, short ones, and ones longer than max_title_len (truncated by the negative pad)
        "tokenized_title": [rng.integers(1, 5000, int(n)).tolist() for n in rng.integers(0, 46, n_news)],
        "tokenized_abstract": [rng.integers(1, 5000, int(n)).tolist() for n in rng.integers(0, 80, n_news)],
        "category_class": rng.integers(1, 19, n_news).tolist(),
        "subcategory_class": rng.integers(1, 200, n_news).tolist(),
        "sentiment_class": rng.integers(0, 3, n_news).tolist(),
        "sentiment_score": rng.normal(size=n_news).astype(np.float64).tolist(),
    }
    news["tokenized_title"][0] = list(range(1, L_TITLE + 1))          # exactly max_title_len
    behaviors = []
    for i in range(n_imp):
        h = int(rng.integers(1, 70))                                  # some histories exceed max_history_len = 50
        c = int(rng.integers(2, 25))
        labels = np.zeros(c, np.int64)
        labels[rng.choice(c, size=int(rng.integers(1, max(2, c // 4 + 1))), replace=False)] = 1
        behaviors.append({"uid": f"U{int(rng.integers(1, 10**6))}", "user": int(rng.integers(0, 50000)),
                          "history": rng.integers(0, n_news, h).tolist(), "candidates": rng.integers(0, n_news, c).tolist(),
                          "labels": labels.tolist()})
    return news, behaviors


def to_frames(news, behaviors):
    idx = [f"N{n}" for n in news["nid"]]
    df = pd.DataFrame({k: v for k, v in news.items() if k != "nid"}, index=idx)
    bhv = pd.DataFrame([{**b, "history": [idx[r] for r in b["history"]], "candidates": [idx[r] for r in b["candidates"]]}
                        for b in behaviors])
    return df, bhv, {name: r for r, name in enumerate(idx)}


def flatten(prefix, batch, out):
    for k, v in batch.items():
        if isinstance(v, dict):
            flatten(f"{prefix}{k}.", v, out)
        else:
            out[f"{prefix}{k}"] = v.numpy()


def main():
    news, behaviors = make_tables()
    df, bhv, row_of = to_frames(news, behaviors)
    attrs = ["title", "abstract", "category", "subcategory", "sentiment_class", "sentiment_score"]
    collate = DatasetCollate(dataset_attributes=attrs, use_plm=False, tokenizer=None, max_title_len=L_TITLE,
                             max_abstract_len=L_ABS, concatenate_inputs=False)
    out = {"meta": np.array([L_TITLE, L_ABS, 50, 4])}
    for k, v in news.items():
        if k.startswith("tokenized"):
            out[f"news.{k}.flat"] = np.array([t for row in v for t in row], np.int64)
            out[f"news.{k}.len"] = np.array([len(row) for row in v], np.int64)
        else:
            out[f"news.{k}"] = np.array(v)
    for split, ds in (("test", RecommendationDatasetTest(df, bhv, max_history_len=50)),
                      ("train", RecommendationDatasetTrain(df, bhv, max_history_len=50, neg_sampling_ratio=4))):
        np.random.seed(1234)                                          # _sample_candidates draws from the global numpy RNG
        items = [ds[i] for i in range(len(ds))]
        ref = collate(items)                                          # the reference's own collate
        # the same impressions as table-row lists, for the restatement (and for DeviceCollate in the GPU tests)
        rows = [(u, ui, np.array([row_of[n] for n in h.index]), np.array([row_of[n] for n in c.index]), lab)
                for (u, ui, h, c, lab) in items]
        mine = CO.collate(news, rows, L_TITLE, L_ABS)
        flat_ref, flat_mine = {}, {}
        flatten("", ref, flat_ref)
        flatten("", mine, flat_mine)
        assert set(flat_ref) == set(flat_mine), (sorted(flat_ref), sorted(flat_mine))
        for k in flat_ref:
            assert flat_ref[k].dtype == flat_mine[k].dtype and np.array_equal(flat_ref[k], flat_mine[k]), (split, k)
            out[f"{split}.batch.{k}"] = flat_ref[k]
        out[f"{split}.hist_rows"] = np.concatenate([r[2] for r in rows])
        out[f"{split}.hist_len"] = np.array([len(r[2]) for r in rows])
        out[f"{split}.cand_rows"] = np.concatenate([r[3] for r in rows])
        out[f"{split}.cand_len"] = np.array([len(r[3]) for r in rows])
        out[f"{split}.labels"] = np.concatenate([r[4] for r in rows])
        out[f"{split}.user_ids"] = np.concatenate([r[0] for r in rows])
        out[f"{split}.user_idx"] = np.concatenate([r[1] for r in rows])
        print(f"[{split}] reference DatasetCollate == collate_oracle.collate on {len(items)} impressions, "
              f"{len(out[f'{split}.hist_rows'])} history rows, {len(out[f'{split}.cand_rows'])} candidates")
    path = os.path.join(GOLD, "collate_ref.npz")
    golden_save(path, **out)
    print("wrote" if "--check" not in sys.argv else "checked", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
