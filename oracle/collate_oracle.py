"""CPU restatement of the reference collate for pre-tokenised (non-PLM) news -- TEST INFRASTRUCTURE,
imported only by ``tests/`` and ``oracle/make_collate_golden.py``.

``DatasetCollate.__call__`` (``newsreclib/data/components/rec_dataset.py:148-168``), ``_tokenize_embeddings``
(``:170-178``, the very same ``F.pad`` call), the scalar columns of ``_tokenize_df`` (``:189-285``) and
``_make_batch_asignees`` (``:289-293``) restated over plain Python lists (table-row indices instead of pandas
``.loc`` on news ids).

Pinned: ``oracle/make_collate_golden.py`` loads the reference's ``rec_dataset.py`` UNMODIFIED (only the
``mind_dataframe`` base-class import, which drags in ``omegaconf`` / ``hydra``, is satisfied by a stand-in), runs
its own ``RecommendationDatasetTest`` / ``RecommendationDatasetTrain`` + ``DatasetCollate`` on a synthetic news
table, asserts that this restatement reproduces every tensor bit for bit, and stores inputs and reference outputs
in ``tests/golden/collate_ref.npz`` (checked by ``tests/test_oracle_cpu.py`` and, for the device-side collate,
``tests/test_gpu_reference_goldens.py``).  The reference's own tests hold no vectors for this path."""
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def tokenize_embeddings(text: List[List[int]], max_len: int) -> torch.Tensor:
    """``rec_dataset.py:170-178`` verbatim semantics (negative pad truncates)."""
    text_padded = [F.pad(torch.tensor(item, dtype=torch.long), (0, max_len - len(item)), "constant", 0) for item in text]
    return torch.vstack(text_padded).long()


def make_batch_asignees(items: Sequence[Sequence]) -> torch.Tensor:
    """``rec_dataset.py:289-293``."""
    sizes = torch.tensor([len(x) for x in items])
    return torch.repeat_interleave(torch.arange(len(items)), sizes)


def collate(news: Dict[str, list], batch, max_title_len: int, max_abstract_len=None) -> Dict:
    """``news``: column -> list indexed by table row (``tokenized_title``, ``tokenized_abstract``,
    ``nid``, ``category_class``, ``subcategory_class``, ``sentiment_class``, ``sentiment_score``)."""
    user_ids, user_idx, histories, candidates, labels = zip(*batch)

    def tokenize_df(rows):
        out = {"news_ids": torch.tensor([news["nid"][r] for r in rows]).long(),
               "title": tokenize_embeddings([news["tokenized_title"][r] for r in rows], max_title_len)}
        if "tokenized_abstract" in news:
            out["abstract"] = tokenize_embeddings([news["tokenized_abstract"][r] for r in rows], max_abstract_len)
        out["category"] = torch.tensor([news["category_class"][r] for r in rows]).long()
        out["subcategory"] = torch.tensor([news["subcategory_class"][r] for r in rows]).long()
        if "sentiment_class" in news:
            out["sentiment"] = torch.tensor([news["sentiment_class"][r] for r in rows]).long()
            out["sentiment_score"] = torch.tensor([news["sentiment_score"][r] for r in rows]).float()
        return out

    return {
        "batch_hist": make_batch_asignees(histories), "batch_cand": make_batch_asignees(candidates),
        "x_hist": tokenize_df([r for h in histories for r in h]),
        "x_cand": tokenize_df([r for c in candidates for r in c]),
        "labels": torch.from_numpy(np.concatenate(labels)).float(),
        "user_ids": torch.from_numpy(np.concatenate(user_ids)).long(),
        "user_idx": torch.from_numpy(np.concatenate(user_idx)).long(),
    }
