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
        if criterion == "cross_entropy_loss":
            return CrossEntropySoft()
        raise ValueError(f"Loss not defined on the sm_100a path: {criterion} "
                         "(the hot path implements cross_entropy_loss, configs/model/nrms.yaml:6)")

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
