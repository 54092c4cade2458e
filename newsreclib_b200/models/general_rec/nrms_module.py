"""Drop-in ``NRMSModule`` (reference ``newsreclib/models/general_rec/nrms_module.py:19-535``):
same constructor kwargs (so ``configs/model/nrms.yaml`` instantiates it by switching
``_target_``), same ``state_dict`` keys, ``forward(batch) -> [B, Cmax]`` fp32 scores and the
same 11-tuple from ``model_step``.  Everything between the batch and the scores runs on the
sm_100a path; there is no ATen fallback for the encoders, the scorer or the loss."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch

from ..components.encoders.news.text import PLM
from .two_tower import TwoTowerRecommender
from ..components.encoders.news.news import NewsEncoder
from ..components.encoders.news.text import MHSAAddAtt
from ..components.encoders.user.nrms import UserEncoder
from ..components.layers.click_predictor import DotProduct


class NRMSModule(TwoTowerRecommender):
    def __init__(
        self,
        dataset_attributes: List[str],
        attributes2encode: List[str],
        outputs: Dict[str, List[str]],
        dual_loss_training: bool,
        dual_loss_coef: Optional[float],
        loss: str,
        late_fusion: bool,
        temperature: Optional[float],
        use_plm: bool,
        pretrained_embeddings_path: Optional[str],
        plm_model: Optional[str],
        frozen_layers: Optional[List[int]],
        embed_dim: int,
        num_heads: int,
        query_dim: int,
        dropout_probability: float,
        top_k_list: List[int],
        num_categ_classes: int,
        num_sent_classes: int,
        save_recs: bool,
        recs_fpath: Optional[str],
        optimizer,
        scheduler,
        pretrained_embeddings: Optional[torch.Tensor] = None,
        transformer_impl: str = "native",
    ) -> None:
        super().__init__(outputs=outputs, optimizer=optimizer, scheduler=scheduler)
        self.num_categ_classes = num_categ_classes + 1
        self.num_sent_classes = num_sent_classes + 1
        if save_recs:
            assert isinstance(recs_fpath, str)
        self.save_recs, self.recs_fpath = save_recs, recs_fpath
        self._init_loss(loss, dual_loss_training, dual_loss_coef)
        if use_plm:
            # nrms_module.py:143-157: HF transformer + the MHSA / additive head (sm_100a)
            assert isinstance(plm_model, (str, torch.nn.Module))
            text_encoder = PLM(plm_model=plm_model, frozen_layers=frozen_layers, embed_dim=embed_dim,
                               use_mhsa=True, apply_reduce_dim=False, reduced_embed_dim=None,
                               num_heads=num_heads, query_dim=query_dim, dropout_probability=dropout_probability,
                               transformer_impl=transformer_impl)
        else:
            if pretrained_embeddings is None:
                assert isinstance(pretrained_embeddings_path, str)
                pretrained_embeddings = self._init_embedding(filepath=pretrained_embeddings_path)
            # RNG consumption order matches the reference __init__ (title MHA, title additive,
            # user MHA, user additive; nrms_module.py:128-171) so seed_everything gives equal inits
            text_encoder = MHSAAddAtt(pretrained_embeddings=pretrained_embeddings, embed_dim=embed_dim,
                                      num_heads=num_heads, query_dim=query_dim,
                                      dropout_probability=dropout_probability)
        self.news_encoder = NewsEncoder(
            dataset_attributes=dataset_attributes, attributes2encode=attributes2encode,
            concatenate_inputs=False, text_encoder=text_encoder, category_encoder=None, entity_encoder=None,
            combine_vectors=False, combine_type=None, input_dim=None, query_dim=None, output_dim=None)
        self.late_fusion = late_fusion
        if not late_fusion:
            self.user_encoder = UserEncoder(news_embed_dim=embed_dim, num_heads=num_heads, query_dim=query_dim)
        self.click_predictor = DotProduct()
        self.use_plm = use_plm
        # model_step as one autograd node on the fused step (ops.NrmsStepFn); NRL_FUSED_MODEL_STEP=0 keeps the per-op
        # Functions.  grad_targets: set by ModuleTrainer to the views of its flat gradient buffer (direct accumulation)
        self.fused_model_step = os.environ.get("NRL_FUSED_MODEL_STEP", "1") != "0"
        self.grad_targets = None
        self.top_k_list = list(top_k_list)
        self.training_step_outputs = {key: [] for key in self.step_outputs["train"]}
        self.val_step_outputs = {key: [] for key in self.step_outputs["val"]}
        self.test_step_outputs = {key: [] for key in self.step_outputs["test"]}

    def _fused_step_inputs(self, batch):
        if not self.fused_model_step or self.use_plm or self.loss_name != "cross_entropy_loss":
            return None
        enc = getattr(self.news_encoder, "text_encoders", None)
        if enc is None or list(enc.keys()) != ["title"] or not isinstance(enc["title"], MHSAAddAtt) \
                or getattr(self.news_encoder, "encode_category", False):
            return None
        if not self.late_fusion and self.user_encoder.attention_axis != "reference":
            return None
        title = enc["title"]
        if not self.late_fusion and self.user_encoder.precision != title.precision:
            return None
        x_h, x_c = batch["x_hist"].get("title"), batch["x_cand"].get("title")
        if not (torch.is_tensor(x_h) and torch.is_tensor(x_c) and x_h.dim() == 2 and x_h.shape[1:] == x_c.shape[1:]):
            return None
        mha, add = title.multihead_attention, title.additive_attention
        params = [title.embedding_layer.weight, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight,
                  mha.out_proj.bias, add.linear.weight, add.linear.bias, add.query]
        if self.late_fusion:
            params += [None] * 7
        else:
            u = self.user_encoder
            params += [u.multihead_attention.in_proj_weight, u.multihead_attention.in_proj_bias,
                       u.multihead_attention.out_proj.weight, u.multihead_attention.out_proj.bias,
                       u.additive_attention.linear.weight, u.additive_attention.linear.bias, u.additive_attention.query]
        training = self.training and title.dropout.p > 0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0  # CPU RNG, no device sync
        targets = None
        if self.grad_targets is not None and torch.is_grad_enabled():
            targets = [None if p is None else p.grad for p in params] if isinstance(self.grad_targets, str) \
                else list(self.grad_targets)
            if any(t is None for t, p in zip(targets, params) if p is not None):
                targets = None  # a parameter without a preallocated .grad: hand the gradients to autograd
        return params, (title.num_heads, bool(self.late_fusion), float(title.dropout.p), bool(training), seed,
                        title.precision, targets)
