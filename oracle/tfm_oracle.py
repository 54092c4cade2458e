"""TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; the
product path never imports it).

CPU restatement of the transformer inside the reference's PLM news encoder:
``self.plm_model(**text)[0]`` at ``newsreclib/models/components/encoders/news/text.py:92`` with
``plm_model = AutoModel.from_pretrained("roberta-base")`` (``:67``).  The algorithm lives in a THIRD-PARTY dependency
of the reference -- HuggingFace ``transformers`` (``modeling_roberta.py``: ``RobertaEmbeddings``,
``RobertaSelfAttention``, ``RobertaSelfOutput``, ``RobertaIntermediate``, ``RobertaOutput``; the reference's
environment pins transformers 4.x, this image carries 5.5.0 with the same mathematics).  The restatement below is
pinned against that library run here: ``oracle/make_tfm_golden.py`` executes the real ``RobertaModel`` (through the
reference's own ``PLM`` constructor / freezing code) and stores its outputs and gradients in
``tests/golden/tfm_*.npz``; ``tests/test_oracle_cpu.py`` checks this file against those fixtures.

Parameters are addressed by HF ``state_dict`` names (``embeddings.word_embeddings.weight``,
``encoder.layer.0.attention.self.query.weight`` ...).  Dropout is replayed from explicit keep masks so that the CUDA
path's own Philox masks can be fed in (``masks`` argument)."""
import math

import torch
import torch.nn.functional as F


def position_ids(input_ids, pad_idx):
    """RobertaEmbeddings.create_position_ids_from_input_ids"""
    m = input_ids.ne(pad_idx).int()
    return (torch.cumsum(m, dim=1).type_as(m) * m).long() + pad_idx


def _drop(x, keep, p):
    if keep is None or p == 0.0:
        return x
    return x * keep.to(x.dtype) / (1.0 - p)


def embeddings(input_ids, P, pad_idx, eps, keep=None, p=0.0, bert=False):
    """RobertaEmbeddings.forward: (word + token_type[0]) + position -> LayerNorm -> dropout.  ``bert=True``:
    BertEmbeddings.forward (modeling_bert.py) -- position ids 0..T-1 and no padding row in the position table."""
    D = P["embeddings.word_embeddings.weight"].shape[1]
    if bert:
        e = F.embedding(input_ids, P["embeddings.word_embeddings.weight"], padding_idx=pad_idx)
        e = e + P["embeddings.token_type_embeddings.weight"][0]
        e = e + P["embeddings.position_embeddings.weight"][: input_ids.shape[1]][None]
        e = F.layer_norm(e, (D,), P["embeddings.LayerNorm.weight"], P["embeddings.LayerNorm.bias"], eps)
        return _drop(e, keep, p)
    # both tables are nn.Embedding(padding_idx=pad_token_id): row pad_idx is read but never receives a gradient
    e = F.embedding(input_ids, P["embeddings.word_embeddings.weight"], padding_idx=pad_idx)
    e = e + P["embeddings.token_type_embeddings.weight"][0]
    e = e + F.embedding(position_ids(input_ids, pad_idx), P["embeddings.position_embeddings.weight"], padding_idx=pad_idx)
    e = F.layer_norm(e, (D,), P["embeddings.LayerNorm.weight"], P["embeddings.LayerNorm.bias"], eps)
    return _drop(e, keep, p)


def layer(x, key_mask, P, pre, heads, eps, masks=None, l=0, p_hidden=0.0, p_attn=0.0):
    """RobertaLayer.forward on x [N, T, D]; key_mask [N, T] bool (True = the key takes part)."""
    masks = masks or {}
    N, T, D = x.shape
    dh = D // heads

    def lin(t, name):
        return t @ P[pre + name + ".weight"].T + P[pre + name + ".bias"]

    def split(t):
        return t.view(N, T, heads, dh).permute(0, 2, 1, 3)
    q, k, v = split(lin(x, "attention.self.query")), split(lin(x, "attention.self.key")), split(lin(x, "attention.self.value"))
    s = q @ k.transpose(-1, -2) / math.sqrt(dh)
    s = s.masked_fill(~key_mask[:, None, None, :], float("-inf"))
    a = torch.softmax(s, dim=-1)
    a = _drop(a, masks.get(("attn", l)), p_attn)
    ctx = (a @ v).permute(0, 2, 1, 3).reshape(N, T, D)
    t1 = _drop(lin(ctx, "attention.output.dense"), masks.get(("attn_out", l)), p_hidden)
    h1 = F.layer_norm(t1 + x, (D,), P[pre + "attention.output.LayerNorm.weight"], P[pre + "attention.output.LayerNorm.bias"], eps)
    u = F.gelu(lin(h1, "intermediate.dense"))
    t2 = _drop(lin(u, "output.dense"), masks.get(("out", l)), p_hidden)
    return F.layer_norm(t2 + h1, (D,), P[pre + "output.LayerNorm.weight"], P[pre + "output.LayerNorm.bias"], eps)


def encoder(input_ids, attention_mask, P, heads, num_layers, pad_idx=1, eps=1e-5, masks=None, p_hidden=0.0, p_attn=0.0,
            bert=False):
    """RobertaModel.forward(...)[0]: last hidden state [N, T, D].  ``masks``: {"embed": keep [N, T, D],
    ("attn_out", l) / ("out", l): keep [N, T, D], ("attn", l): keep [N, heads, T, T]} or None (eval)."""
    masks = masks or {}
    x = embeddings(input_ids, P, pad_idx, eps, masks.get("embed"), p_hidden, bert=bert)
    km = torch.ones_like(input_ids, dtype=torch.bool) if attention_mask is None else attention_mask.bool()
    for l in range(num_layers):
        x = layer(x, km, P, f"encoder.layer.{l}.", heads, eps, masks, l, p_hidden, p_attn)
    return x
