"""NAML ``UserEncoder`` (reference ``newsreclib/models/components/encoders/user/naml.py:7-31``):
additive pooling of the dense clicked-news vectors (zero-padded rows included)."""
import torch
import torch.nn as nn

from newsreclib_b200.models.components.layers.attention import AdditiveAttention


class UserEncoder(nn.Module):
    def __init__(self, news_embed_dim: int, query_dim: int) -> None:
        super().__init__()
        self.additive_attention = AdditiveAttention(input_dim=news_embed_dim, query_dim=query_dim)

    def forward(self, hist_news_vector: torch.Tensor) -> torch.Tensor:
        """``[B, Hmax, D]`` -> ``[B, D]``."""
        return self.additive_attention(hist_news_vector)
