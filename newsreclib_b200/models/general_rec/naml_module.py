"""Drop-in ``NAMLModule`` (reference ``newsreclib/models/general_rec/naml_module.py:20-566``): same
constructor kwargs (``configs/model/naml.yaml`` instantiates it by switching ``_target_``), same
``state_dict`` keys (title and abstract alias ONE ``CNNAddAtt`` instance, ``news.py:68-77``),
``forward(batch) -> [B, Cmax]`` and the 11-tuple ``model_step``.  News encoder = CNN text encoder on
title and abstract + category encoder, combined by additive attention over the three views; user
encoder = additive pooling; all on the sm_100a path."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ..components.encoders.news.category import LinearEncoder
from ..components.encoders.news.news import NewsEncoder
from ..components.encoders.news.text import PLM, CNNAddAtt
from ..components.encoders.user.naml import UserEncoder
from ..components.layers.click_predictor import DotProduct
from .two_tower import TwoTowerRecommender


class NAMLModule(TwoTowerRecommender):
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
        text_embed_dim: int,
        num_heads: int,
        num_filters: Optional[int],
        window_size: Optional[int],
        query_dim: int,
        categ_embed_dim: int,
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
        # construction order = the reference's (naml_module.py:128-205): text encoder, category
        # encoder, view combiner, user encoder -> same RNG consumption under seed_everything
        if not use_plm:
            assert isinstance(num_filters, int) and isinstance(window_size, int)
            if pretrained_embeddings is None:
                assert isinstance(pretrained_embeddings_path, str)
                pretrained_embeddings = self._init_embedding(filepath=pretrained_embeddings_path)
            text_encoder = CNNAddAtt(pretrained_embeddings=pretrained_embeddings, embed_dim=text_embed_dim,
                                     num_filters=num_filters, window_size=window_size, query_dim=query_dim,
                                     dropout_probability=dropout_probability)
            news_dim = num_filters
        else:
            assert isinstance(plm_model, (str, torch.nn.Module))
            text_encoder = PLM(plm_model=plm_model, frozen_layers=frozen_layers, embed_dim=text_embed_dim,
                               use_mhsa=True, apply_reduce_dim=False, reduced_embed_dim=None, num_heads=num_heads,
                               query_dim=query_dim, dropout_probability=dropout_probability,
                               transformer_impl=transformer_impl)
            news_dim = text_embed_dim
        category_encoder = LinearEncoder(
            pretrained_embeddings=None, from_pretrained=False, freeze_pretrained_emb=False,
            num_categories=self.num_categ_classes, embed_dim=categ_embed_dim, use_dropout=False,
            dropout_probability=None, linear_transform=True, output_dim=news_dim)
        self.news_encoder = NewsEncoder(
            dataset_attributes=dataset_attributes, attributes2encode=attributes2encode, concatenate_inputs=False,
            text_encoder=text_encoder, category_encoder=category_encoder, entity_encoder=None, combine_vectors=True,
            combine_type="add_att", input_dim=news_dim, query_dim=query_dim, output_dim=None)
        self.late_fusion = late_fusion
        if not late_fusion:
            self.user_encoder = UserEncoder(news_embed_dim=news_dim, query_dim=query_dim)
        self.click_predictor = DotProduct()
        self.top_k_list = list(top_k_list)
        self.training_step_outputs = {key: [] for key in self.step_outputs["train"]}
        self.val_step_outputs = {key: [] for key in self.step_outputs["val"]}
        self.test_step_outputs = {key: [] for key in self.step_outputs["test"]}
