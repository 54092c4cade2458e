"""Base class of the drop-in recommender modules: the slice of the reference's
``AbstractRecommneder`` (``newsreclib/models/abstract_recommender.py:14-193``) that the NRMS /
NAML modules rely on — hyper-parameter capture, optimizer construction from Hydra partials,
embedding loading, loss selection and the step-output bookkeeping.

Subclasses ``lightning.LightningModule`` when Lightning is importable (so Hydra + Trainer drive
it exactly like the reference module); in images without Lightning (this one) it degrades to
an ``nn.Module`` with no-op ``log`` / ``save_hyperparameters`` shims so the same class can be
driven by a plain training loop, the tests and ``bench.py``."""
from __future__ import annotations

import inspect
from types import SimpleNamespace
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

try:  # pragma: no cover - depends on the image
    from lightning import LightningModule as _Base

    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    HAVE_LIGHTNING = False

    class _Base(nn.Module):  # minimal stand-in for LightningModule
        def __init__(self) -> None:
            super().__init__()
            self.hparams = SimpleNamespace()

        def save_hyperparameters(self, logger: bool = False, ignore=None) -> None:
            frame = inspect.currentframe().f_back
            while frame is not None and frame.f_code.co_name != "__init__":
                frame = frame.f_back
            # walk up to the outermost __init__ of this object (the concrete subclass)
            outer = frame
            while outer is not None and outer.f_back is not None and \
                    outer.f_back.f_code.co_name == "__init__" and outer.f_back.f_locals.get("self") is self:
                outer = outer.f_back
            args = {k: v for k, v in (outer.f_locals if outer else {}).items()
                    if k not in ("self", "__class__") and not k.startswith("_")}
            self.hparams = SimpleNamespace(**args)

        @property
        def device(self) -> torch.device:
            p = next(self.parameters(), None)
            return p.device if p is not None else torch.device("cpu")

        def log(self, *a, **k) -> None:
            pass

        def log_dict(self, *a, **k) -> None:
            pass


class CrossEntropySoft(nn.Module):
    """``torch.nn.CrossEntropyLoss()`` on float targets, evaluated by the CUDA kernel when the
    ragged layout (labels + offsets) is supplied, see ``NRMSModule.model_step``."""

    def forward(self, scores: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor:
        from .. import ops
        B, Cn = scores.shape
        off = torch.arange(0, (B + 1) * Cn, Cn, dtype=torch.int32, device=scores.device)
        return ops.CESoftFn.apply(scores, y_true.reshape(-1).float(), off)


class SupConLoss(nn.Module):
    """The reference's ``SupConLoss`` (``models/components/losses.py:6-40``, a pytorch-metric-learning subclass fed with
    the score matrix and the positive / negative index tuples of ``nrms_module.py:290-307``) on the ragged layout:
    positives are the real candidates with a non-zero label, negatives the real candidates with label 0.  The reference
    always builds it with the default temperature (``abstract_recommender.py:117-120``)."""

    def __init__(self, temperature: float = 0.1) -> None:
        super().__init__()
        self.temperature = temperature

    def forward(self, scores: torch.Tensor, labels: torch.Tensor, cand_off: torch.Tensor) -> torch.Tensor:
        from .. import ops
        if self.temperature != ops.SupConFn.TEMPERATURE:
            raise ValueError("the reference never passes a temperature to SupConLoss(): only the default 0.1 exists")
        return ops.SupConFn.apply(scores, labels, cand_off, None)


class AbstractRecommneder(_Base):  # sic: the reference's class name (abstract_recommender.py:14)
    def __init__(self, outputs: Dict[str, List[str]], optimizer, scheduler) -> None:
        super().__init__()
        self.save_hyperparameters(logger=False)
        self.step_outputs = {stage: list(keys) for stage, keys in outputs.items()}

    # -- optimizer (abstract_recommender.py:89-108) ------------------------------------
    def configure_optimizers(self) -> Dict[str, Any]:
        optimizer = self.hparams.optimizer(params=self.parameters())
        if getattr(self.hparams, "scheduler", None) is not None:
            scheduler = self.hparams.scheduler(optimizer=optimizer)
            return {"optimizer": optimizer,
                    "lr_scheduler": {"scheduler": scheduler, "monitor": "valid/loss", "interval": "epoch",
                                     "frequency": 1}}
        return {"optimizer": optimizer}

    # -- helpers (abstract_recommender.py:110-157) ---------------------------------------
    def _init_embedding(self, filepath: str) -> torch.Tensor:
        return torch.from_numpy(np.load(filepath)).float()

    def _get_loss(self, criterion: str) -> Union[Callable, Tuple[Callable, Callable]]:
        # abstract_recommender.py:113-124
        if criterion == "cross_entropy_loss":
            return CrossEntropySoft()
        if criterion == "sup_con_loss":
            return SupConLoss()
        if criterion == "dual_loss":
            return CrossEntropySoft(), SupConLoss()
        raise ValueError(f"Loss not defined: {criterion}")

    def _init_loss(self, loss: str, dual_loss_training: bool, dual_loss_coef) -> None:
        """``nrms_module.py:113-119`` / ``naml_module.py`` (same lines)."""
        self.loss_name, self.dual_loss_coef = loss, dual_loss_coef
        if not dual_loss_training:
            self.criterion = self._get_loss(loss)
            if isinstance(self.criterion, tuple):
                raise ValueError("loss='dual_loss' needs dual_loss_training=True (the reference would fail in model_step)")
        else:
            assert isinstance(dual_loss_coef, float)
            assert loss == "dual_loss"
            self.ce_criterion, self.scl_criterion = self._get_loss(loss)

    def _loss(self, scores: torch.Tensor, labels: torch.Tensor, cand_off: torch.Tensor) -> torch.Tensor:
        """``nrms_module.py:286-328`` on the ragged layout (labels ``[N_c]`` + candidate offsets)."""
        from .. import ops
        labels = labels.float().contiguous()
        if self.loss_name == "cross_entropy_loss":
            return ops.CESoftFn.apply(scores, labels, cand_off)
        if self.loss_name == "sup_con_loss":
            return self.criterion(scores, labels, cand_off)
        return ops.SupConFn.apply(scores, labels, cand_off, float(self.dual_loss_coef))  # both kernels, one Function

    def _get_recommendations(self, user_ids: torch.Tensor, news_ids: torch.Tensor, scores: torch.Tensor,
                             cand_news_size: torch.Tensor) -> Dict[str, Dict[str, float]]:
        """``abstract_recommender.py:150-181``: ``{"U<user id>": {"N<news id>": score, ...}, ...}`` -- one entry per user,
        later impressions of the same user add to (and overwrite within) that user's dictionary."""
        owner = torch.repeat_interleave(user_ids.detach().cpu(), cand_news_size.detach().cpu()).tolist()
        recs: Dict[str, Dict[str, float]] = {}
        for u, n, s in zip(owner, news_ids.detach().cpu().tolist(), scores.detach().cpu().tolist()):
            recs.setdefault(f"U{u}", {})[f"N{n}"] = s
        return recs

    def _save_recommendations(self, recommendations: Dict[str, Dict[str, float]], fpath: str) -> None:
        """``abstract_recommender.py:183-185``: one JSON file."""
        import json
        with open(fpath, "w") as f:
            json.dump(recommendations, f)

    def _collect_model_outputs(self, vector: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        # rows are concatenated in impression order: identical to boolean-mask indexing
        return vector[mask]

    def _collect_step_outputs(self, outputs_dict, local_vars):
        for key in outputs_dict.keys():
            outputs_dict[key].append(local_vars.get(key, []))
        return outputs_dict

    def _gather_step_outputs(self, outputs_dict, key: str) -> torch.Tensor:
        if key not in outputs_dict.keys():
            raise AttributeError(f"{key} not in {outputs_dict}")
        return torch.cat([o for o in outputs_dict[key]])

    def _clear_epoch_outputs(self, outputs_dict):
        for key in outputs_dict.keys():
            outputs_dict[key].clear()
        return outputs_dict


AbstractRecommender = AbstractRecommneder
