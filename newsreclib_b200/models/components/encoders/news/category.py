"""``LinearEncoder`` category encoder (reference
``newsreclib/models/components/encoders/news/category.py:8-82``): same 9 constructor kwargs,
attribute names (``embedding_layer``, ``dropout``, ``linear``) and ``state_dict`` keys; forward and
backward (embedding gather -> optional dropout -> Linear -> ReLU) on the sm_100a path."""
from typing import Optional

import torch
import torch.nn as nn

from newsreclib_b200 import ops


class LinearEncoder(nn.Module):
    def __init__(self, pretrained_embeddings: Optional[torch.Tensor], from_pretrained: bool,
                 freeze_pretrained_emb: bool, num_categories: int, embed_dim: Optional[int], use_dropout: bool,
                 dropout_probability: Optional[float], linear_transform: bool, output_dim: Optional[int]) -> None:
        super().__init__()
        if from_pretrained:
            assert isinstance(pretrained_embeddings, torch.Tensor)
            self.embedding_layer = nn.Embedding.from_pretrained(
                embeddings=pretrained_embeddings, freeze=freeze_pretrained_emb, padding_idx=0)
        else:
            assert isinstance(embed_dim, int) and embed_dim > 0
            self.embedding_layer = nn.Embedding(num_embeddings=num_categories, embedding_dim=embed_dim,
                                                padding_idx=0)
        self.use_dropout = use_dropout
        if self.use_dropout:
            if not isinstance(dropout_probability, float):
                raise ValueError(
                    f"Expected keyword argument `dropout_probability` to be a `float` but got {dropout_probability}")
            self.dropout = nn.Dropout(p=dropout_probability)
        self.linear_transform = linear_transform
        if self.linear_transform:
            assert isinstance(output_dim, int)
            self.linear = nn.Linear(in_features=self.embedding_layer.embedding_dim, out_features=output_dim)
        else:
            raise NotImplementedError("LinearEncoder without linear_transform is outside the NAML hot path "
                                      "(configs/model/naml.yaml builds it with linear_transform=True)")
        self.precision = ops.PREC_BF16X3

    def forward(self, category: torch.Tensor) -> torch.Tensor:
        """category: int64 ``[N]`` -> ``[N, output_dim]``."""
        p = float(self.dropout.p) if self.use_dropout else 0.0
        training = self.training and p > 0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0
        return ops.LinearEncoderFn.apply(category.contiguous(), self.embedding_layer.weight, self.linear.weight,
                                         self.linear.bias, p, training, seed, self.precision)
