"""``NewsEncoder`` with the reference's 11 constructor kwargs
(``newsreclib/models/components/encoders/news/news.py:38-51``).  The NRMS configuration
(one text attribute, ``combine_vectors=False``) returns the single text vector; the NAML
configuration stacks the title / abstract / category views and combines them with the
sm_100a ``AdditiveAttention`` (``news.py:162-163``), forward and backward.  ``torch.stack`` of the
``[N, D]`` view vectors is the only ATen op in between (a 3 x N x D copy, as in the reference)."""
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from newsreclib_b200.models.components.layers.attention import AdditiveAttention


class NewsEncoder(nn.Module):
    def __init__(self, dataset_attributes: List[str], attributes2encode: List[str], concatenate_inputs: bool,
                 text_encoder: Optional[nn.Module], category_encoder: Optional[nn.Module],
                 entity_encoder: Optional[nn.Module], combine_vectors: bool, combine_type: Optional[str],
                 input_dim: Optional[int], query_dim: Optional[int], output_dim: Optional[int]) -> None:
        super().__init__()
        assert len(dataset_attributes) > 0
        self.concatenate_inputs = concatenate_inputs
        wanted = set(dataset_attributes) & set(attributes2encode)
        self.encode_text = bool({"title", "abstract"} & set(attributes2encode))
        self.encode_category = bool({"category", "subcategory"} & set(attributes2encode))
        self.encode_entity = bool({"title_entities", "abstract_entities"} & set(attributes2encode))
        if self.encode_text:
            assert isinstance(text_encoder, nn.Module)
            names = ["text"] if concatenate_inputs else [n for n in ("title", "abstract") if n in wanted]
            # one module instance registered under every text attribute (shared weights, as in
            # the reference: both key prefixes then alias the same tensors in the state_dict)
            self.text_encoders = nn.ModuleDict({n: text_encoder for n in names})
        if self.encode_category:
            assert isinstance(category_encoder, nn.Module)
            self.category_encoders = nn.ModuleDict(
                {n: category_encoder for n in ("category", "subcategory") if n in wanted})
        if self.encode_entity:
            raise NotImplementedError("entity encoders are outside the NRMS/NAML hot path")
        if combine_vectors:
            assert isinstance(combine_type, str)
            self.combine_type = combine_type
            if combine_type == "add_att":
                assert isinstance(input_dim, int) and input_dim > 0
                assert isinstance(query_dim, int) and query_dim > 0
                self.combine_layer = AdditiveAttention(input_dim=input_dim, query_dim=query_dim)
            elif combine_type == "linear":
                assert isinstance(input_dim, int) and input_dim > 0
                assert isinstance(output_dim, int) and output_dim > 0
                self.combine_layer = nn.Linear(in_features=input_dim, out_features=output_dim)
            elif combine_type == "concat":
                self.combine_layer = lambda vectors: torch.cat(vectors, dim=1)
            else:
                raise ValueError(
                    f"Expected keyword argument `combine_type` to be in [`add_att`, `linear`, `concat`] but got {combine_type}.")

    def forward(self, news: Dict[str, torch.Tensor]) -> torch.Tensor:
        vectors = []
        if self.encode_text:
            vectors += [enc(news[name]) for name, enc in self.text_encoders.items()]
        if self.encode_category:
            vectors += [enc(news[name]) for name, enc in self.category_encoders.items()]
        return self._combine(vectors)

    def forward_pair(self, news_a: Dict[str, torch.Tensor], news_b: Dict[str, torch.Tensor]):
        """``(self(news_a), self(news_b))``; text encoders that can share work between the two calls (the PLM: one pass
        through the transformer, ``PLM.forward_pair``) do so, everything else is called twice."""
        va, vb = [], []
        if self.encode_text:
            for name, enc in self.text_encoders.items():
                a, b = enc.forward_pair(news_a[name], news_b[name]) if hasattr(enc, "forward_pair") else \
                    (enc(news_a[name]), enc(news_b[name]))
                va.append(a); vb.append(b)
        if self.encode_category:
            for name, enc in self.category_encoders.items():
                va.append(enc(news_a[name])); vb.append(enc(news_b[name]))
        return self._combine(va), self._combine(vb)

    def _combine(self, vectors) -> torch.Tensor:
        if len(vectors) == 1:
            return vectors[0]
        if self.combine_type == "add_att":
            return self.combine_layer(torch.stack(vectors, dim=1))
        if self.combine_type == "linear":
            return self.combine_layer(torch.cat(vectors, dim=-1))
        return self.combine_layer(vectors)
