"""Stand-ins that let the reference's OWN ``NRMSModule`` / ``NAMLModule`` files be imported unmodified.

THIS IS TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE (same rule as the rest of ``oracle/``: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it).

``newsreclib/models/general_rec/nrms_module.py`` imports ``lightning``, ``torchmetrics``, ``torch_geometric`` and
(through ``newsreclib.models.components.losses``) ``pytorch_metric_learning``.  None of them is in this image and none
can be installed (no network).  Only ONE of those imports takes part in the arithmetic of ``forward`` /
``model_step`` with the cross-entropy loss: ``torch_geometric.utils.to_dense_batch`` (call sites
``nrms_module.py:233,237,277-284``).  ``install()`` registers:

* ``lightning.LightningModule`` -> ``torch.nn.Module`` + ``save_hyperparameters`` (constructor arguments ->
  ``self.hparams``), a ``device`` property and no-op ``log`` / ``log_dict``;
* ``torchmetrics`` metric classes and ``newsreclib.metrics.*`` -> inert objects (constructed in ``__init__``, never
  called by ``forward`` / ``model_step``);
* ``pytorch_metric_learning`` -> ``oracle/pml_standins.py`` (restatement of the few published 2.2.0 pieces under the
  reference's ``SupConLoss``), so the reference's own ``newsreclib/models/components/losses.py`` loads unmodified and
  ``loss="sup_con_loss"`` / ``"dual_loss"`` run through the reference's own ``model_step`` (``nrms_module.py:290-328``);
* ``torch_geometric.utils.to_dense_batch`` -> ``oracle.nrms_oracle.to_dense_batch``, the restatement of the published
  PyG 2.3.0 algorithm -- third-party, absent, "parity unpinned" for that one function.

The reference package itself comes from ``ref_root`` (``/root/reference`` in the build container, ``baseline/_ref``
-- the offline install made by ``baseline/install_ref.py`` -- on the GPU box) and is not touched.
"""
from __future__ import annotations

import inspect
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASELINE_REF = os.path.join(ROOT, "baseline", "_ref")


class _HParams(dict):
    __getattr__ = dict.__getitem__


class LightningModule(torch.nn.Module):
    def save_hyperparameters(self, *args, **kwargs):
        hp = _HParams()
        frame = inspect.currentframe().f_back
        while frame is not None:  # every __init__ of the class chain that is on the stack contributes its arguments
            if frame.f_code.co_name == "__init__" and frame.f_locals.get("self") is self:
                for k, v in frame.f_locals.items():
                    if k not in ("self", "__class__", "args", "kwargs"):
                        hp.setdefault(k, v)
            frame = frame.f_back
        self.hparams = hp

    @property
    def device(self):
        return next(self.parameters()).device if any(True for _ in self.parameters()) else torch.device("cpu")

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass


class Inert(torch.nn.Module):
    """Placeholder for metric / loss objects the timed and checked code paths never call."""

    def __init__(self, *a, **k):
        super().__init__()

    def clone(self, prefix=None):
        return Inert()

    def add_metrics(self, *a, **k):
        pass

    def reset(self):
        pass


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def reference_root() -> str:
    """Where the unmodified reference package lives: the source tree in the build container, else the offline
    install under ``baseline/_ref`` (which travels to the GPU box)."""
    if os.path.isdir("/root/reference/newsreclib"):
        return "/root/reference"
    if os.path.isdir(os.path.join(BASELINE_REF, "newsreclib")):
        return BASELINE_REF
    raise ImportError("the reference package is neither at /root/reference nor installed under baseline/_ref "
                      "(run `python baseline/install_ref.py` in the build container)")


def install(ref_root: str | None = None) -> str:
    """Register the stand-ins and put the reference package on ``sys.path``.  Returns the root used."""
    from oracle import nrms_oracle as O

    ref_root = ref_root or reference_root()
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    _module("lightning", LightningModule=LightningModule)
    _module("torch_geometric")
    _module("torch_geometric.utils", to_dense_batch=O.to_dense_batch)
    _module("torchmetrics", MetricCollection=Inert, MeanMetric=Inert, MinMetric=Inert)
    _module("torchmetrics.classification", AUROC=Inert)
    _module("torchmetrics.retrieval", RetrievalMRR=Inert, RetrievalNormalizedDCG=Inert)
    _module("newsreclib.metrics.diversity", Diversity=Inert)
    _module("newsreclib.metrics.personalization", Personalization=Inert)
    from oracle import pml_standins
    pml_standins.install()  # the reference's own losses.py is imported on top of it
    return ref_root
