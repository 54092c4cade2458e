"""NRMS ``UserEncoder`` (``newsreclib/models/components/encoders/user/nrms.py:7-41``): same
constructor and ``state_dict`` keys; forward/backward on the sm_100a path.

``attention_axis="reference"`` (default) reproduces the reference exactly: its
``nn.MultiheadAttention`` is ``batch_first=False`` and receives ``[B, Hmax, E]`` without a
permute, so self-attention runs ACROSS THE B IMPRESSIONS at each history position.
``attention_axis="history"`` attends along the clicked-news history instead."""
import torch
import torch.nn as nn

from newsreclib_b200 import ops
from newsreclib_b200.models.components.layers.attention import AdditiveAttention


class UserEncoder(nn.Module):
    def __init__(self, news_embed_dim: int, num_heads: int, query_dim: int,
                 attention_axis: str = "reference") -> None:
        super().__init__()
        if attention_axis not in ("reference", "history"):
            raise ValueError(f"attention_axis must be 'reference' or 'history', got {attention_axis}")
        self.multihead_attention = nn.MultiheadAttention(embed_dim=news_embed_dim, num_heads=num_heads)
        self.additive_attention = AdditiveAttention(input_dim=news_embed_dim, query_dim=query_dim)
        self.num_heads = num_heads
        self.attention_axis = attention_axis
        self.precision = ops.PREC_BF16X3

    def forward(self, hist_news_vector: torch.Tensor) -> torch.Tensor:
        """``[B, Hmax, E]`` dense (zero-padded) history vectors -> ``[B, E]`` user vectors."""
        mha, add = self.multihead_attention, self.additive_attention
        return ops.UserEncoderFn.apply(
            hist_news_vector, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias,
            add.linear.weight, add.linear.bias, add.query, self.num_heads,
            0 if self.attention_axis == "reference" else 1, self.precision)
