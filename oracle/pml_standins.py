"""Stand-in for the pieces of pytorch-metric-learning 2.2.0 (``setup.py:26`` of the reference) that the reference's
``newsreclib/models/components/losses.py`` builds its ``SupConLoss`` on.

THIS IS TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE (same rule as the rest of ``oracle/``).

pytorch-metric-learning is a third-party dependency that is neither under ``/root/reference`` nor in this image (no
network): its algorithm is restated here from the published 2.2.0 source and is "parity unpinned" by the reference
(no test or golden vector of the reference touches it).  What IS pinned: with this module registered as
``pytorch_metric_learning`` the reference's own ``losses.py`` (``compute_loss`` / ``_compute_loss`` overrides,
``:12-40``) and its own ``model_step`` (index tuples, ``nrms_module.py:290-328``) run UNMODIFIED, and
``oracle/make_module_golden.py`` asserts that ``oracle.nrms_oracle.sup_con_loss`` reproduces their value and gradient.

Restated (pytorch_metric_learning 2.2.0):
* ``losses.BaseMetricLossFunction.forward`` -> ``compute_loss`` -> ``reducer`` (``losses/base_metric_loss_function.py``);
* ``losses.GenericPairLoss.mat_based_loss``: ``pos_mask[a1, p] = 1; neg_mask[a2, n] = 1`` (``losses/generic_pair_loss.py``);
* ``zero_losses()``: ``{"loss": {"losses": 0, "indices": None, "reduction_type": "already_reduced"}}``;
* ``SupConLoss.get_default_reducer() = AvgNonZeroReducer()`` = ``ThresholdReducer(low=0)``: the mean of the elements
  ``> 0``, ``sum(embeddings * 0)`` when there is none or when the loss dict is the zero loss (``reducers/*.py``);
* ``utils.common_functions``: ``small_val = finfo.tiny``, ``neg_inf = finfo.min``, ``torch_arange_from_size``;
* ``utils.loss_and_miner_utils.logsumexp``: masked entries filled with ``finfo.min``, optional appended zero, rows
  without any kept entry give 0.
"""
from __future__ import annotations

import sys
import types

import torch


# ---------------------------------------------------------------- utils.common_functions
def small_val(dtype):
    return torch.finfo(dtype).tiny


def neg_inf(dtype):
    return torch.finfo(dtype).min


def torch_arange_from_size(x, size_dim=0):
    return torch.arange(x.size(size_dim), device=x.device)


# ---------------------------------------------------------------- utils.loss_and_miner_utils
def logsumexp(x, keep_mask=None, add_one=True, dim=1):
    if keep_mask is not None:
        x = x.masked_fill(~keep_mask, neg_inf(x.dtype))
    if add_one:
        zeros = torch.zeros(x.size(dim - 1), dtype=x.dtype, device=x.device).unsqueeze(dim)
        x = torch.cat([x, zeros], dim=dim)
    output = torch.logsumexp(x, dim=dim, keepdim=True)
    if keep_mask is not None:
        output = output.masked_fill(~torch.any(keep_mask, dim=dim, keepdim=True), 0)
    return output


# ---------------------------------------------------------------- reducers.AvgNonZeroReducer
def avg_non_zero_reduce(loss_dict, embeddings):
    assert len(loss_dict) == 1
    info = next(iter(loss_dict.values()))
    losses, kind = info["losses"], info["reduction_type"]
    if (not torch.is_tensor(losses)) and losses == 0:
        return torch.sum(embeddings * 0)
    if kind == "already_reduced":
        assert losses.ndim == 0 or len(losses) == 1
        return losses
    assert kind == "element"
    keep = losses > 0
    if int(keep.sum()) >= 1:
        return torch.mean(losses[keep])
    return torch.sum(embeddings * 0)


# ---------------------------------------------------------------- losses.SupConLoss (the base the reference subclasses)
class SupConLoss(torch.nn.Module):
    def __init__(self, temperature=0.1, **kwargs):
        super().__init__()
        self.temperature = temperature

    def add_to_recordable_attributes(self, name=None, list_of_names=None, is_stat=False):
        pass

    def zero_losses(self):
        return {"loss": {"losses": 0, "indices": None, "reduction_type": "already_reduced"}}

    def mat_based_loss(self, mat, indices_tuple):
        a1, p, a2, n = indices_tuple
        pos_mask, neg_mask = torch.zeros_like(mat), torch.zeros_like(mat)
        pos_mask[a1, p] = 1
        neg_mask[a2, n] = 1
        return self._compute_loss(mat, pos_mask, neg_mask)

    loss_method = mat_based_loss  # GenericPairLoss(mat_based_loss=True)

    def forward(self, embeddings, labels=None, indices_tuple=None, ref_emb=None, ref_labels=None):
        if ref_emb is None:
            ref_emb, ref_labels = embeddings, labels
        loss_dict = self.compute_loss(embeddings, labels, indices_tuple, ref_emb, ref_labels)
        return avg_non_zero_reduce(loss_dict, embeddings)


def install() -> None:
    """Register the stand-in as ``pytorch_metric_learning`` (+ the three sub-modules ``losses.py:1-3`` imports)."""
    pkg = types.ModuleType("pytorch_metric_learning")
    losses = types.ModuleType("pytorch_metric_learning.losses")
    losses.SupConLoss = SupConLoss
    utils = types.ModuleType("pytorch_metric_learning.utils")
    c_f = types.ModuleType("pytorch_metric_learning.utils.common_functions")
    c_f.small_val, c_f.neg_inf, c_f.torch_arange_from_size = small_val, neg_inf, torch_arange_from_size
    lmu = types.ModuleType("pytorch_metric_learning.utils.loss_and_miner_utils")
    lmu.logsumexp = logsumexp
    pkg.losses, pkg.utils = losses, utils
    utils.common_functions, utils.loss_and_miner_utils = c_f, lmu
    for m in (pkg, losses, utils, c_f, lmu):
        sys.modules[m.__name__] = m
