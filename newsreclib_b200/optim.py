"""``torch.optim.Adam`` on the sm_100a Adam kernel, for the drop-in modules' ``configure_optimizers``.

The reference builds its optimizer from a Hydra partial (``configs/model/nrms.yaml:49-52``: ``_target_: torch.optim.Adam,
lr: 0.0001``; ``abstract_recommender.py:89-108``).  ``newsreclib_b200.optim.Adam`` is a ``torch.optim.Optimizer`` with the
same constructor arguments for the options those configs use (lr, betas, eps; ``weight_decay`` / ``amsgrad`` /
``maximize`` other than their defaults raise) whose ``step`` is one ``nrl_adam_step`` launch per parameter: one pass over
p / g / m / v instead of torch's multi-pass ``foreach`` update -- on the 84 MB embedding table that is the whole cost of the
optimizer.  Same arithmetic as ``torch/optim/adam.py::_single_tensor_adam`` (dense: rows with a zero gradient still see
their moments decay); ``state_dict`` uses torch's key names (``step``, ``exp_avg``, ``exp_avg_sq``) so optimizer checkpoints
move between the two.
"""
from __future__ import annotations

import torch

from . import ops


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 amsgrad: bool = False, maximize: bool = False, zero_grad_in_step: bool = False) -> None:
        """``zero_grad_in_step=True`` clears each ``.grad`` in the same pass that consumes it (then ``zero_grad()`` has
        nothing left to do; use with ``zero_grad(set_to_none=False)`` or none at all)."""
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError(f"invalid Adam hyper-parameters lr={lr} betas={betas} eps={eps}")
        if weight_decay != 0.0 or amsgrad or maximize:
            raise NotImplementedError("newsreclib_b200.optim.Adam covers the options of the reference configs "
                                      "(lr, betas, eps); weight_decay / amsgrad / maximize are not built")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0.0, amsgrad=False, maximize=False))
        self.zero_grad_in_step = bool(zero_grad_in_step)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse:
                    raise RuntimeError("newsreclib_b200.optim.Adam does not support sparse gradients")
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("newsreclib_b200.optim.Adam needs contiguous fp32 CUDA parameters "
                                       "(there is no CPU path)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] = int(st["step"]) + 1
                ops.adam_step(p.data, g if g.is_contiguous() else g.contiguous(), st["exp_avg"], st["exp_avg_sq"],
                              st["step"], group["lr"], b1, b2, group["eps"],
                              zero_grad=self.zero_grad_in_step and g.is_contiguous())
        return loss
