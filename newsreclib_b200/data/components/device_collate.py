"""Device-side collate: the reference builds every batch on the host with one
pandas ``.loc`` per impression and one ``F.pad`` + ``vstack`` per title
(``newsreclib/data/components/rec_dataset.py:39-55,148-178,189-293``, ``num_workers=0``), which caps
the input pipeline far below what the sm_100a encoders consume.  Here the news table is tokenised and
padded ONCE (same pad / truncate rule as ``_tokenize_embeddings``, ``rec_dataset.py:170-178``), kept
resident in HBM, and a batch is assembled by index: the per-impression history / candidate row lists
are concatenated on the host (a few KB), uploaded, and ``nrl_gather_rows`` pulls the token-id rows and
per-news scalars.  The result is the reference's ``RecommendationBatch`` layout, bit for bit."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ... import ops
from .batch import RecommendationBatch


def pad_token_lists(token_lists: Sequence[Sequence[int]], max_len: int) -> np.ndarray:
    """``_tokenize_embeddings`` (``rec_dataset.py:170-178``) for a fixed ``max_len``: right-pad with 0,
    or truncate (the reference's negative ``F.pad`` drops the tail) -> int64 ``[n, max_len]``."""
    out = np.zeros((len(token_lists), max_len), dtype=np.int64)
    for i, item in enumerate(token_lists):
        k = min(len(item), max_len)
        if k:
            out[i, :k] = np.asarray(item[:k], dtype=np.int64)
    return out


class DeviceNewsTable:
    """All per-news attributes the collate emits, as dense device tensors indexed by table row."""

    def __init__(self, news_ids: np.ndarray, title: np.ndarray, category: np.ndarray, subcategory: np.ndarray,
                 abstract: Optional[np.ndarray] = None, sentiment: Optional[np.ndarray] = None,
                 sentiment_score: Optional[np.ndarray] = None, device="cuda") -> None:
        dev = torch.device(device)
        self.num_news = int(title.shape[0])
        self.cols: Dict[str, torch.Tensor] = {
            "news_ids": torch.as_tensor(news_ids, dtype=torch.int64).to(dev).contiguous(),
            "title": torch.as_tensor(title, dtype=torch.int64).to(dev).contiguous(),
            "category": torch.as_tensor(category, dtype=torch.int64).to(dev).contiguous(),
            "subcategory": torch.as_tensor(subcategory, dtype=torch.int64).to(dev).contiguous(),
        }
        if abstract is not None:
            self.cols["abstract"] = torch.as_tensor(abstract, dtype=torch.int64).to(dev).contiguous()
        if sentiment is not None:
            self.cols["sentiment"] = torch.as_tensor(sentiment, dtype=torch.int64).to(dev).contiguous()
            self.cols["sentiment_score"] = torch.as_tensor(sentiment_score, dtype=torch.float32).to(dev).contiguous()
        self.device = dev

    @classmethod
    def from_token_lists(cls, news_ids, tokenized_titles, category, subcategory, max_title_len: int,
                         tokenized_abstracts=None, max_abstract_len: Optional[int] = None, sentiment=None,
                         sentiment_score=None, device="cuda") -> "DeviceNewsTable":
        abstract = None
        if tokenized_abstracts is not None:
            assert isinstance(max_abstract_len, int) and max_abstract_len > 0
            abstract = pad_token_lists(tokenized_abstracts, max_abstract_len)
        return cls(np.asarray(news_ids), pad_token_lists(tokenized_titles, max_title_len), np.asarray(category),
                   np.asarray(subcategory), abstract, None if sentiment is None else np.asarray(sentiment),
                   None if sentiment_score is None else np.asarray(sentiment_score), device)

    def gather(self, rows: torch.Tensor) -> Dict[str, torch.Tensor]:
        """``_tokenize_df`` (``rec_dataset.py:189-285``) by index: every column gathered at ``rows``."""
        return {k: ops.gather_rows(v, rows) for k, v in self.cols.items()}


class DeviceCollate:
    """``DatasetCollate.__call__`` (``rec_dataset.py:148-168``) over ``DeviceNewsTable`` rows.

    A sample is ``(user_id, user_idx, history_rows, candidate_rows, labels)`` with the two row lists
    indexing the table (what ``self.news.loc[...]`` resolves to in ``rec_dataset.py:51-52``)."""

    def __init__(self, table: DeviceNewsTable) -> None:
        self.table = table

    def __call__(self, batch: Sequence[Tuple]) -> RecommendationBatch:
        user_ids, user_idx, histories, candidates, labels = zip(*batch)
        hist_rows = np.concatenate([np.asarray(h, dtype=np.int64) for h in histories])
        cand_rows = np.concatenate([np.asarray(c, dtype=np.int64) for c in candidates])
        for rows in (hist_rows, cand_rows):
            if rows.size and (rows.min() < 0 or rows.max() >= self.table.num_news):
                raise KeyError("news row index outside the table")  # pandas .loc raises KeyError too
        dev = self.table.device
        sizes_h = np.array([len(h) for h in histories], dtype=np.int64)
        sizes_c = np.array([len(c) for c in candidates], dtype=np.int64)
        # _make_batch_asignees (rec_dataset.py:289-293): sorted segment ids
        batch_hist = torch.from_numpy(np.repeat(np.arange(len(batch), dtype=np.int64), sizes_h)).to(dev)
        batch_cand = torch.from_numpy(np.repeat(np.arange(len(batch), dtype=np.int64), sizes_c)).to(dev)
        x_hist = self.table.gather(torch.from_numpy(hist_rows).to(dev))
        x_cand = self.table.gather(torch.from_numpy(cand_rows).to(dev))
        # dense_widths: (Hmax, Cmax), known here on the host -- the modules then need no device sync to size the dense
        # history / candidate tensors (the reference's to_dense_batch reads them back from the device)
        return RecommendationBatch(
            dense_widths=(int(sizes_h.max()), int(sizes_c.max())),
            batch_hist=batch_hist, batch_cand=batch_cand, x_hist=x_hist, x_cand=x_cand,
            labels=torch.from_numpy(np.concatenate([np.asarray(l) for l in labels])).float().to(dev),
            user_ids=torch.from_numpy(np.concatenate([np.atleast_1d(u) for u in user_ids])).long().to(dev),
            user_idx=torch.from_numpy(np.concatenate([np.atleast_1d(u) for u in user_idx])).long().to(dev))
