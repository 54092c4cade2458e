"""Text encoders of the hot path.  ``MHSAAddAtt`` keeps the reference's constructor, attribute
names and ``state_dict`` keys (``newsreclib/models/components/encoders/news/text.py:179-236``):
stock ``nn.Embedding`` / ``nn.MultiheadAttention`` / ``AdditiveAttention`` / ``nn.Dropout``
sub-modules are the PARAMETER CONTAINERS (so checkpoints, seeded initial weights, optimizer
groups and DDP bucketing match the reference by construction); ``forward`` hands their tensors
to the sm_100a encoder (gather -> MHSA on tcgen05 -> additive pooling), forward and backward."""
import torch
import torch.nn as nn

from newsreclib_b200 import ops
from newsreclib_b200.models.components.layers.attention import AdditiveAttention


class MHSAAddAtt(nn.Module):
    def __init__(self, pretrained_embeddings: torch.Tensor, embed_dim: int, num_heads: int,
                 query_dim: int, dropout_probability: float) -> None:
        super().__init__()
        if not isinstance(dropout_probability, float):
            raise ValueError(
                f"Expected keyword argument `dropout_probability` to be a `float` but got {dropout_probability}")
        self.embedding_layer = nn.Embedding.from_pretrained(
            torch.as_tensor(pretrained_embeddings, dtype=torch.float32), freeze=False, padding_idx=0)
        self.multihead_attention = nn.MultiheadAttention(embed_dim=embed_dim, num_heads=num_heads)
        self.additive_attention = AdditiveAttention(input_dim=embed_dim, query_dim=query_dim)
        self.dropout = nn.Dropout(dropout_probability)
        self.num_heads = num_heads
        self.precision = ops.PREC_BF16X3

    def forward(self, text: torch.Tensor) -> torch.Tensor:
        """text: int64 ``[N, L]`` token ids (0 = pad, embedded with table row 0 like the
        reference) -> ``[N, E]``."""
        mha, add = self.multihead_attention, self.additive_attention
        training = self.training and self.dropout.p > 0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0  # CPU RNG, no device sync
        return ops.NewsEncoderFn.apply(
            text.contiguous(), self.embedding_layer.weight, mha.in_proj_weight, mha.in_proj_bias,
            mha.out_proj.weight, mha.out_proj.bias, add.linear.weight, add.linear.bias, add.query,
            self.num_heads, float(self.dropout.p), training, seed, self.precision)
