"""``DotProduct`` click predictor (``newsreclib/models/components/layers/click_predictor.py:5-11``)."""
import torch
import torch.nn as nn

from newsreclib_b200 import ops


class DotProduct(nn.Module):
    def __init__(self) -> None:
        super().__init__()

    def forward(self, user_vec: torch.Tensor, cand_news_vector: torch.Tensor) -> torch.Tensor:
        """user_vec ``[B, 1, E]``, cand_news_vector ``[B, E, C]`` (the reference's dense call,
        ``nrms_module.py:251-253``) -> ``[B, C]``."""
        B, E, Cn = cand_news_vector.shape
        cand = cand_news_vector.permute(0, 2, 1).contiguous().reshape(B * Cn, E)
        off = torch.arange(0, (B + 1) * Cn, Cn, dtype=torch.int32, device=cand.device)
        return ops.ScoreFn.apply(user_vec.reshape(B, E), cand, off, B, Cn)

    @staticmethod
    def ragged(user: torch.Tensor, cand: torch.Tensor, cand_off: torch.Tensor, B: int, Cmax: int) -> torch.Tensor:
        """Same scores straight from the ragged candidate vectors (no dense ``[B, Cmax, E]``)."""
        return ops.ScoreFn.apply(user, cand, cand_off, B, Cmax)
