"""Batch layout of the hot path's input: mirrors ``newsreclib/data/components/batch.py:6-32``
(field names and meanings are the boundary; ``DatasetCollate`` of the reference produces it)."""
from typing import Any, Dict, TypedDict

import torch


class RecommendationBatch(TypedDict, total=False):
    """Ragged PyG-style recommendation batch.

    batch_hist / batch_cand: sorted int64 segment ids (impression index of every history /
    candidate row); x_hist / x_cand: per-row news features (``title`` int64 ``[N, L]`` …);
    labels: float32 click labels per candidate row; user_ids / user_idx: int64 ``[B]``.
    """

    batch_hist: torch.Tensor
    batch_cand: torch.Tensor
    x_hist: Dict[str, Any]
    x_cand: Dict[str, Any]
    labels: torch.Tensor
    user_ids: torch.Tensor
    user_idx: torch.Tensor
    # optional, not in the reference: (Hmax, Cmax) as host integers when the collate knows them (DeviceCollate)
    dense_widths: tuple


class NewsBatch(TypedDict):
    """News-only batch (``batch.py:35-51`` of the reference); not consumed by the hot path."""

    news: Dict[str, Any]
    labels: torch.Tensor
