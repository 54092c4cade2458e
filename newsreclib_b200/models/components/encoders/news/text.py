"""Text encoders of the hot path (``MHSAAddAtt`` for NRMS, ``CNNAddAtt`` for NAML, ``PLM`` =
HF transformer + the sm_100a head).  ``MHSAAddAtt`` keeps the reference's constructor, attribute
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


class CNNAddAtt(nn.Module):
    """NAML text encoder (reference ``text.py:112-176``): same constructor, attribute names
    (``embedding_layer``, ``cnn``, ``additive_attention``, ``dropout``) and ``state_dict`` keys.
    forward = gather -> dropout -> Conv2d(1, F, (w, E), pad ((w-1)/2, 0)) as ONE K = w*E tcgen05 GEMM
    over an im2col of the gathered rows -> ReLU -> dropout -> additive pooling."""

    def __init__(self, pretrained_embeddings: torch.Tensor, embed_dim: int, num_filters: int, window_size: int,
                 query_dim: int, dropout_probability: float) -> None:
        super().__init__()
        if not isinstance(dropout_probability, float):
            raise ValueError(
                f"Expected keyword argument `dropout_probability` to be a `float` but got {dropout_probability}")
        self.embedding_layer = nn.Embedding.from_pretrained(
            torch.as_tensor(pretrained_embeddings, dtype=torch.float32), freeze=False, padding_idx=0)
        self.cnn = nn.Conv2d(in_channels=1, out_channels=num_filters, kernel_size=(window_size, embed_dim),
                             padding=(int((window_size - 1) / 2), 0))
        self.additive_attention = AdditiveAttention(input_dim=num_filters, query_dim=query_dim)
        self.dropout = nn.Dropout(dropout_probability)
        self.window_size = window_size
        self.precision = ops.PREC_BF16X3

    def forward(self, text: torch.Tensor) -> torch.Tensor:
        """text: int64 ``[N, L]`` -> ``[N, num_filters]``."""
        add = self.additive_attention
        training = self.training and self.dropout.p > 0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0
        return ops.CnnEncoderFn.apply(
            text.contiguous(), self.embedding_layer.weight, self.cnn.weight, self.cnn.bias, add.linear.weight,
            add.linear.bias, add.query, self.window_size, float(self.dropout.p), training, seed, self.precision)


class PLM(nn.Module):
    """PLM text encoder (reference ``text.py:15-109``), ``use_mhsa=True`` form used by NRMS-PLM /
    NAML-PLM: the transformer followed by the sm_100a head -- dropout -> MHSA over dim 0 of
    ``[N, T, E]`` (the reference passes batch-first states to a ``batch_first=False`` attention,
    ``text.py:96``, so attention runs across the N news of the call at each token position;
    ``attention_axis="tokens"`` attends along the tokens instead) -> dropout -> additive pooling over
    the T tokens (pad tokens included).

    ``transformer_impl="native"`` (default): the HF ``RobertaModel`` is the PARAMETER
    CONTAINER (``state_dict`` keys, ``from_pretrained`` checkpoints, the reference's name-based freezing
    ``text.py:70-73`` all unchanged) and ``self.plm_model(**text)[0]`` (``text.py:92``) runs on the sm_100a
    encoder (``ops.TfmEncoderFn``: embeddings, every layer's projections on tcgen05, key-padding-masked attention,
    LayerNorm / GELU / the three dropouts, forward and backward).  Architectures the kernels do not cover
    (anything but RoBERTa / XLM-R / BERT post-LN layers with head dim 64, erf-GELU, absolute positions) raise;
    ``transformer_impl="hf"`` keeps the third-party torch module on the path instead."""

    def __init__(self, plm_model, frozen_layers, embed_dim: int, use_mhsa: bool, apply_reduce_dim: bool,
                 reduced_embed_dim, num_heads, query_dim, dropout_probability: float,
                 attention_axis: str = "reference", transformer_impl: str = "native") -> None:
        super().__init__()
        if not isinstance(plm_model, (str, nn.Module)):
            raise ValueError(f"Expected keyword argument `plm_model` to be a `str` but got {plm_model}")
        if not isinstance(dropout_probability, float):
            raise ValueError(
                f"Expected keyword argument `dropout_probability` to be a `float` but got {dropout_probability}")
        if attention_axis not in ("reference", "tokens"):
            raise ValueError(f"attention_axis must be 'reference' or 'tokens', got {attention_axis}")
        self.use_mhsa = use_mhsa
        self.apply_reduce_dim = apply_reduce_dim
        if isinstance(plm_model, nn.Module):  # an already-built transformer (tests, offline images)
            self.plm_model = plm_model
        else:
            from transformers import AutoModel
            self.plm_model = AutoModel.from_pretrained(plm_model)
        for name, param in self.plm_model.base_model.named_parameters():  # text.py:70-73
            for layer in (frozen_layers or []):
                if "layer." + str(layer) + "." in name:
                    param.requires_grad = False
        if self.use_mhsa:
            assert isinstance(num_heads, int) and num_heads > 0
            self.multihead_attention = nn.MultiheadAttention(embed_dim=embed_dim, num_heads=num_heads)
            self.additive_attention = AdditiveAttention(input_dim=embed_dim, query_dim=query_dim)
            self.dropout = nn.Dropout(p=dropout_probability)
        if self.apply_reduce_dim:
            raise NotImplementedError("apply_reduce_dim (MANNeR / MINER configurations) is outside the NRMS/NAML path")
        self.num_heads = num_heads
        self.attention_axis = attention_axis
        self.precision = ops.PREC_BF16X3
        if transformer_impl not in ("native", "hf"):
            raise ValueError(f"transformer_impl must be 'native' or 'hf', got {transformer_impl}")
        self.transformer_impl = transformer_impl
        self._tfm_state = None
        if transformer_impl == "native":
            self._tfm_state = self._native_state()

    def _native_state(self) -> "ops.TfmState":
        cfg = getattr(self.plm_model, "config", None)
        why = None
        if cfg is None or getattr(cfg, "model_type", None) not in ("roberta", "xlm-roberta", "bert"):
            why = f"model_type {getattr(cfg, 'model_type', None)!r} (RoBERTa / XLM-R / BERT only)"
        elif cfg.hidden_size != 64 * cfg.num_attention_heads:
            why = f"head dim {cfg.hidden_size // cfg.num_attention_heads} (64 only)"
        elif cfg.hidden_act != "gelu":
            why = f"hidden_act {cfg.hidden_act!r} (erf 'gelu' only)"
        elif getattr(cfg, "position_embedding_type", "absolute") != "absolute":
            why = "relative position embeddings"
        elif cfg.hidden_size > 1024 or cfg.intermediate_size % 16:
            why = f"hidden {cfg.hidden_size} / intermediate {cfg.intermediate_size}"
        elif getattr(cfg, "is_decoder", False) or getattr(cfg, "add_cross_attention", False):
            why = "decoder / cross-attention configuration"
        if why is not None:
            raise ValueError(f"the sm_100a transformer does not cover this PLM: {why}; pass transformer_impl='hf' to keep "
                             f"the HF torch module on the path")
        return ops.TfmState(cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size, cfg.num_hidden_layers,
                            cfg.vocab_size, cfg.max_position_embeddings, cfg.pad_token_id, cfg.layer_norm_eps,
                            cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob,
                            position_mode=1 if cfg.model_type == "bert" else 0)

    def transformer_parameters(self):
        """The tensors ``ops.TfmEncoderFn`` takes, in ``_lib.TFM_EMBED_FIELDS`` + per-layer ``TFM_LAYER_FIELDS`` order."""
        m = self.plm_model
        e = m.embeddings
        ps = [e.word_embeddings.weight, e.position_embeddings.weight, e.token_type_embeddings.weight[0],
              e.LayerNorm.weight, e.LayerNorm.bias]
        for lyr in m.encoder.layer:
            a, so = lyr.attention.self, lyr.attention.output
            ps += [a.query.weight, a.query.bias, a.key.weight, a.key.bias, a.value.weight, a.value.bias,
                   so.dense.weight, so.dense.bias, so.LayerNorm.weight, so.LayerNorm.bias,
                   lyr.intermediate.dense.weight, lyr.intermediate.dense.bias,
                   lyr.output.dense.weight, lyr.output.dense.bias, lyr.output.LayerNorm.weight, lyr.output.LayerNorm.bias]
        return ps

    def transformer(self, text) -> torch.Tensor:
        """``self.plm_model(**text)[0]``: ``[N, T, D]`` last hidden states."""
        if self.transformer_impl == "hf":
            return self.plm_model(**text)[0]
        extra = set(text.keys()) - {"input_ids", "attention_mask"}
        if extra:
            raise ValueError(f"the sm_100a transformer takes input_ids / attention_mask only, got {sorted(extra)} "
                             f"(the reference tokenises with return_token_type_ids=False, rec_dataset.py:181)")
        ids = text["input_ids"]
        training = self.plm_model.training
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0
        return ops.TfmEncoderFn.apply(ids, text.get("attention_mask"), self._tfm_state, training, seed, self.precision,
                                      *self.transformer_parameters())

    def forward_pair(self, text_a, text_b):
        """``(self(text_a), self(text_b))`` with ONE pass through the transformer: the hidden state of a token does not
        depend on how much padding follows it (padding is masked out of every softmax and keeps position id
        ``pad_token_id``), so the two calls of a training step (clicked history, candidates: ``nrms_module.py:231,235``)
        run as one batch padded to the longer of the two, and each part is cut back to ITS OWN padded length before
        the head -- which, as in the reference, does look at the padding positions and across the news of its call."""
        if self.transformer_impl != "native" or not self.use_mhsa:
            return self(text_a), self(text_b)
        ia, ib = text_a["input_ids"], text_b["input_ids"]
        ma, mb = text_a.get("attention_mask"), text_b.get("attention_mask")
        ma = torch.ones_like(ia) if ma is None else ma
        mb = torch.ones_like(ib) if mb is None else mb
        Ta, Tb = ia.shape[1], ib.shape[1]
        T = max(Ta, Tb)
        pad = int(self.plm_model.config.pad_token_id)

        def widen(ids, mask):
            if ids.shape[1] == T:
                return ids, mask
            extra = T - ids.shape[1]
            return (torch.nn.functional.pad(ids, (0, extra), value=pad), torch.nn.functional.pad(mask, (0, extra), value=0))
        ia2, ma2 = widen(ia, ma)
        ib2, mb2 = widen(ib, mb)
        states = self.transformer({"input_ids": torch.cat([ia2, ib2], 0), "attention_mask": torch.cat([ma2, mb2], 0)})
        na = ia.shape[0]
        return self.head(states[:na, :Ta].contiguous()), self.head(states[na:, :Tb].contiguous())

    def head(self, states: torch.Tensor) -> torch.Tensor:
        """``[N, T, E]`` last hidden states -> ``[N, E]`` on the sm_100a path."""
        mha, add = self.multihead_attention, self.additive_attention
        training = self.training and self.dropout.p > 0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0
        return ops.PlmHeadFn.apply(
            states.float(), mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias,
            add.linear.weight, add.linear.bias, add.query, self.num_heads,
            0 if self.attention_axis == "reference" else 1, float(self.dropout.p), training, seed, self.precision)

    def forward(self, text) -> torch.Tensor:
        if self.use_mhsa:
            return self.head(self.transformer(text))
        return self.transformer(text)[:, 0, :]
