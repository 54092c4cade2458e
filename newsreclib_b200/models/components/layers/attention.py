"""``AdditiveAttention`` with the reference's constructor, parameters and ``state_dict`` keys
(``newsreclib/models/components/layers/attention.py:6-42``); forward runs the sm_100a path
(tcgen05 projection with the tanh / query-dot fused in the epilogue, then the softmax pooling
kernel).  Forward-only as a standalone module (NAML inference); inside the NRMS blocks the
same math is differentiated by the fused encoder kernels."""
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
        if torch.is_grad_enabled() and (input_vector.requires_grad or self.query.requires_grad):
            if input_vector.requires_grad:
                raise RuntimeError("standalone AdditiveAttention is forward-only on the sm_100a path; "
                                   "wrap the call in torch.no_grad() (training uses the fused NRMS blocks)")
        with torch.no_grad():
            return ops.additive_attention(input_vector.float(), self.linear.weight, self.linear.bias,
                                          self.query, self.precision)
