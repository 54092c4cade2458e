"""``AdditiveAttention`` with the reference's constructor, parameters and ``state_dict`` keys
(``newsreclib/models/components/layers/attention.py:6-42``); forward and backward run the
sm_100a path (tcgen05 projection with the tanh / query-dot fused in the epilogue, softmax
pooling kernel; backward through softmax, query dot and tanh + the two gradient GEMMs).
Used standalone by the NAML view combiner (``encoders/news/news.py:162-163``) and the NAML user
encoder (``encoders/user/naml.py:27-31``); inside the NRMS blocks the same math is part of the
fused encoder calls."""
import torch
import torch.nn as nn

from newsreclib_b200 import ops


class AdditiveAttention(nn.Module):
    def __init__(self, input_dim: int, query_dim: int) -> None:
        super().__init__()
        for name, val in (("input_dim", input_dim), ("query_dim", query_dim)):
            if not isinstance(val, int):
                raise ValueError(f"Expected keyword argument `{name}` to be an `int` but got {val}")
        # parameter containers under the reference's attribute names (linear.weight [Q, D],
        # linear.bias [Q], query [Q] ~ U(-0.1, 0.1))
        self.linear = nn.Linear(in_features=input_dim, out_features=query_dim)
        self.query = nn.Parameter(torch.empty(query_dim).uniform_(-0.1, 0.1))
        self.precision = ops.PREC_BF16X3

    def forward(self, input_vector: torch.Tensor) -> torch.Tensor:
        """``[G, L, D]`` -> ``[G, D]``: softmax(tanh(xW^T + b) . q) over dim 1, weighted sum of x."""
        return ops.AdditiveFn.apply(input_vector.float(), self.linear.weight, self.linear.bias, self.query,
                                    self.precision)
