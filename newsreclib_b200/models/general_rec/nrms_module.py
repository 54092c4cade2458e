"""Drop-in ``NRMSModule`` (reference ``newsreclib/models/general_rec/nrms_module.py:19-535``):
same constructor kwargs (so ``configs/model/nrms.yaml`` instantiates it by switching
``_target_``), same ``state_dict`` keys, ``forward(batch) -> [B, Cmax]`` fp32 scores and the
same 11-tuple from ``model_step``.  Everything between the batch and the scores runs on the
sm_100a path; there is no ATen fallback for the encoders, the scorer or the loss."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from ... import ops
from ...data.components.batch import RecommendationBatch
from ...metrics import ranking_metrics
from ..abstract_recommender import AbstractRecommneder
from ..components.encoders.news.news import NewsEncoder
from ..components.encoders.news.text import MHSAAddAtt
from ..components.encoders.user.nrms import UserEncoder
from ..components.layers.click_predictor import DotProduct


class NRMSModule(AbstractRecommneder):
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
    ) -> None:
        super().__init__(outputs=outputs, optimizer=optimizer, scheduler=scheduler)
        self.num_categ_classes = num_categ_classes + 1
        self.num_sent_classes = num_sent_classes + 1
        if save_recs:
            assert isinstance(recs_fpath, str)
        if dual_loss_training:
            raise NotImplementedError("dual_loss_training (SupCon) is outside the hot path; "
                                      "configs/model/nrms.yaml uses cross_entropy_loss")
        self.criterion = self._get_loss(loss)
        if use_plm:
            raise NotImplementedError("the PLM text encoder (roberta-base) is a 'next' row (SURVEY.md §8f-3)")
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
        self.top_k_list = list(top_k_list)
        self.training_step_outputs = {key: [] for key in self.step_outputs["train"]}
        self.val_step_outputs = {key: [] for key in self.step_outputs["val"]}
        self.test_step_outputs = {key: [] for key in self.step_outputs["test"]}

    # ------------------------------------------------------------------ layout helpers
    @staticmethod
    def _layout(batch: RecommendationBatch):
        """Offsets and dense widths of the ragged batch.  One host sync for (B, Hmax, Cmax), the
        same information ``to_dense_batch`` fetches with ``batch.max()`` / ``num.max()``."""
        seg_h, seg_c = batch["batch_hist"], batch["batch_cand"]
        B = int(batch["user_idx"].numel()) if "user_idx" in batch else int(seg_c[-1]) + 1
        off_h, off_c = ops.segment_offsets(seg_h, B), ops.segment_offsets(seg_c, B)
        widths = torch.stack([(off_h[1:] - off_h[:-1]).max(), (off_c[1:] - off_c[:-1]).max()]).tolist()
        return B, off_h, off_c, int(widths[0]), int(widths[1])

    # ------------------------------------------------------------------ forward (nrms_module.py:230-255)
    def forward(self, batch: RecommendationBatch) -> torch.Tensor:
        return self._forward_with_layout(batch, self._layout(batch))

    def _forward_with_layout(self, batch, layout) -> torch.Tensor:
        B, off_h, off_c, Hmax, Cmax = layout
        hist_news_vector = self.news_encoder(batch["x_hist"])
        cand_news_vector = self.news_encoder(batch["x_cand"])
        if not self.late_fusion:
            hist_agg = ops.ToDenseFn.apply(hist_news_vector, off_h, B, Hmax)
            user_vector = self.user_encoder(hist_agg)
        else:
            sizes = (off_h[1:] - off_h[:-1]).to(hist_news_vector.dtype)
            hist_agg = ops.ToDenseFn.apply(hist_news_vector, off_h, B, Hmax)
            user_vector = hist_agg.sum(dim=1) / sizes.unsqueeze(-1)
        return DotProduct.ragged(user_vector, cand_news_vector, off_c, B, Cmax)

    # ------------------------------------------------------------------ model_step (nrms_module.py:260-362)
    def model_step(self, batch: RecommendationBatch) -> Tuple[torch.Tensor, ...]:
        layout = self._layout(batch)
        B, off_h, off_c, Hmax, Cmax = layout
        scores = self._forward_with_layout(batch, layout)
        loss = ops.CESoftFn.apply(scores, batch["labels"].float().contiguous(), off_c)
        cand_news_size = (off_c[1:] - off_c[:-1]).long()
        hist_news_size = (off_h[1:] - off_h[:-1]).long()
        mask_cand = torch.arange(Cmax, device=scores.device)[None, :] < cand_news_size[:, None]
        preds = self._collect_model_outputs(scores, mask_cand)
        targets = batch["labels"]                      # ragged order == masked dense order
        target_categories = batch["x_cand"].get("category")
        target_sentiments = batch["x_cand"].get("sentiment")
        hist_categories = batch["x_hist"].get("category")
        hist_sentiments = batch["x_hist"].get("sentiment")
        return (loss, preds, targets, cand_news_size, hist_news_size, target_categories, target_sentiments,
                hist_categories, hist_sentiments, batch["user_ids"] if "user_ids" in batch else None,
                batch["x_cand"].get("news_ids"))

    # ------------------------------------------------------------------ Lightning hooks
    def training_step(self, batch: RecommendationBatch, batch_idx: int):
        loss, preds, targets, cand_news_size, *_ = self.model_step(batch)
        self.log("train/loss", loss, on_step=True, on_epoch=True, prog_bar=True)
        self.training_step_outputs = self._collect_step_outputs(self.training_step_outputs, locals())
        return loss

    def _epoch_metrics(self, outputs, prefix: str) -> Dict[str, torch.Tensor]:
        preds = self._gather_step_outputs(outputs, "preds")
        targets = self._gather_step_outputs(outputs, "targets")
        sizes = self._gather_step_outputs(outputs, "cand_news_size")
        m = {prefix + k: v for k, v in ranking_metrics(preds, targets, sizes, self.top_k_list).items()}
        self.log_dict(m, on_step=False, on_epoch=True, prog_bar=True, sync_dist=True)
        self._clear_epoch_outputs(outputs)
        return m

    def on_train_epoch_end(self):
        return self._epoch_metrics(self.training_step_outputs, "train/")

    def validation_step(self, batch: RecommendationBatch, batch_idx: int):
        loss, preds, targets, cand_news_size, *_ = self.model_step(batch)
        self.log("val/loss", loss, on_step=False, on_epoch=True, prog_bar=True, sync_dist=True)
        self.val_step_outputs = self._collect_step_outputs(self.val_step_outputs, locals())
        return loss

    def on_validation_epoch_end(self):
        return self._epoch_metrics(self.val_step_outputs, "val/")

    def test_step(self, batch: RecommendationBatch, batch_idx: int):
        (loss, preds, targets, cand_news_size, hist_news_size, target_categories, target_sentiments,
         hist_categories, hist_sentiments, user_ids, cand_news_ids) = self.model_step(batch)
        self.log("test/loss", loss, on_step=False, on_epoch=True, prog_bar=True, sync_dist=True)
        self.test_step_outputs = self._collect_step_outputs(self.test_step_outputs, locals())
        return loss

    def on_test_epoch_end(self):
        return self._epoch_metrics(self.test_step_outputs, "test/")
