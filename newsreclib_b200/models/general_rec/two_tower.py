"""Shared body of the two-tower recommenders (NRMS, NAML): the reference's ``forward`` /
``model_step`` / Lightning step hooks are line-for-line identical in ``nrms_module.py:230-535`` and
``naml_module.py:261-566`` (``diff`` of everything from ``forward`` to EOF is empty), so they live
once here.  ``forward`` = news encoder over history and candidates -> ragged->dense -> user encoder
(or the late-fusion mean) -> dot-product scores; everything runs on the sm_100a path."""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from ... import ops
from ...data.components.batch import RecommendationBatch
from ...metrics import aspect_metrics, ranking_metrics
from ..abstract_recommender import AbstractRecommneder
from ..components.layers.click_predictor import DotProduct


class TwoTowerRecommender(AbstractRecommneder):
    late_fusion: bool = False
    # the NRMS / NAML news encoders treat every news of a call independently (attention within one title, pooling per
    # news), so ``news_encoder(x_hist)`` and ``news_encoder(x_cand)`` (nrms_module.py:231,235) give the same vectors
    # as ONE call over the concatenated news: half the launches, twice the rows per GEMM, one embedding-gradient
    # buffer in the backward pass.  Not valid for the PLM head (it attends across the news of a call), which keeps two.
    merge_news_calls: bool = True

    def _encode_hist_and_cand(self, x_hist, x_cand):
        enc = self.news_encoder
        if not hasattr(self, "_per_news_encoder"):
            from ..components.encoders.news.text import PLM
            self._per_news_encoder = not any(isinstance(m, PLM) for m in enc.modules())
        keys = list(getattr(enc, "text_encoders", {}).keys()) + list(getattr(enc, "category_encoders", {}).keys())
        if not (self.merge_news_calls and self._per_news_encoder and keys and all(
                torch.is_tensor(x_hist.get(k)) and torch.is_tensor(x_cand.get(k)) and x_hist[k].shape[1:] == x_cand[k].shape[1:]
                for k in keys)):
            if hasattr(enc, "forward_pair") and self.merge_news_calls:
                return enc.forward_pair(x_hist, x_cand)  # PLM: one pass through the transformer, two through the head
            return enc(x_hist), enc(x_cand)
        n_hist = x_hist[keys[0]].shape[0]
        vec = enc({k: torch.cat([x_hist[k], x_cand[k]], dim=0) for k in keys})
        return vec[:n_hist], vec[n_hist:]

    # ------------------------------------------------------------------ layout helpers
    @staticmethod
    def _layout(batch: RecommendationBatch):
        """Offsets and dense widths of the ragged batch.  (B, Hmax, Cmax) are host integers: ``to_dense_batch`` fetches
        them with a device sync (``batch.max()`` / ``num.max()``); a collate that already knows them on the host --
        ``DeviceCollate``, the synthetic generator -- passes them as ``batch["dense_widths"] = (Hmax, Cmax)`` and the
        step then has no sync at all, so the host keeps running ahead of the GPU."""
        seg_h, seg_c = batch["batch_hist"], batch["batch_cand"]
        B = int(batch["user_idx"].numel()) if "user_idx" in batch else int(seg_c[-1]) + 1
        off_h, off_c = ops.segment_offsets(seg_h, B), ops.segment_offsets(seg_c, B)
        widths = batch.get("dense_widths")
        if widths is None:
            widths = torch.stack([(off_h[1:] - off_h[:-1]).max(), (off_c[1:] - off_c[:-1]).max()]).tolist()
        return B, off_h, off_c, int(widths[0]), int(widths[1])

    # ------------------------------------------------------------------ forward (nrms_module.py:230-255)
    def forward(self, batch: RecommendationBatch) -> torch.Tensor:
        return self._forward_with_layout(batch, self._layout(batch))

    def _forward_with_layout(self, batch, layout) -> torch.Tensor:
        B, off_h, off_c, Hmax, Cmax = layout
        hist_news_vector, cand_news_vector = self._encode_hist_and_cand(batch["x_hist"], batch["x_cand"])
        if not self.late_fusion:
            hist_agg = ops.ToDenseFn.apply(hist_news_vector, off_h, B, Hmax)
            user_vector = self.user_encoder(hist_agg)
        else:
            user_vector = ops.LateFusionFn.apply(hist_news_vector, off_h, B)
        return DotProduct.ragged(user_vector, cand_news_vector, off_c, B, Cmax)

    # ------------------------------------------------------------------ evaluation with cached news vectors
    # validation / test re-encode every candidate of every impression
    # (nrms_module.py:398-535); a news vector only depends on the news (the text encoders attend
    # within one title and pad tokens are unmasked), so in eval mode the table is encoded ONCE and
    # an impression batch becomes two row gathers + the user encoder + the scorer.
    @torch.no_grad()
    def encode_news_table(self, news: Dict[str, torch.Tensor], chunk: int = 16384) -> torch.Tensor:
        """Eval-mode news vectors ``[num_news, D]`` for a whole news table (dict of per-news columns)."""
        from ..components.encoders.news.text import PLM
        if any(isinstance(m, PLM) for m in self.news_encoder.modules()):
            raise NotImplementedError("the PLM head attends ACROSS the news of a call (text.py:96): its vectors "
                                      "depend on the batch composition and cannot be cached per news")
        was_training = self.training
        self.eval()
        n = next(iter(news.values())).shape[0]
        out = [self.news_encoder({k: v[i:i + chunk].contiguous() for k, v in news.items()}) for i in range(0, n, chunk)]
        self.train(was_training)
        return torch.cat(out, dim=0)

    @torch.no_grad()
    def forward_cached(self, news_vectors: torch.Tensor, hist_rows: torch.Tensor, batch_hist: torch.Tensor,
                       cand_rows: torch.Tensor, batch_cand: torch.Tensor, B: int) -> torch.Tensor:
        """Scores ``[B, Cmax]`` from cached news vectors; equals ``forward`` in eval mode."""
        off_h, off_c = ops.segment_offsets(batch_hist, B), ops.segment_offsets(batch_cand, B)
        widths = torch.stack([(off_h[1:] - off_h[:-1]).max(), (off_c[1:] - off_c[:-1]).max()]).tolist()
        Hmax, Cmax = int(widths[0]), int(widths[1])
        hist = ops.gather_rows(news_vectors, hist_rows)
        cand = ops.gather_rows(news_vectors, cand_rows)
        if not self.late_fusion:
            user = self.user_encoder(ops.ToDenseFn.apply(hist, off_h, B, Hmax))
        else:
            user = ops.LateFusionFn.apply(hist, off_h, B)
        return DotProduct.ragged(user, cand, off_c, B, Cmax)

    def _fused_step_inputs(self, batch):
        """Parameters + configuration of ``ops.NrmsStepFn`` when this module is the plain NRMS configuration it covers,
        else ``None`` (subclasses that have a fused step override this)."""
        return None

    # ------------------------------------------------------------------ model_step (nrms_module.py:260-362)
    def model_step(self, batch: RecommendationBatch) -> Tuple[torch.Tensor, ...]:
        layout = self._layout(batch)
        B, off_h, off_c, Hmax, Cmax = layout
        fused = self._fused_step_inputs(batch)
        if fused is not None:
            # NRMS + cross-entropy: the whole differentiable part as ONE autograd node on the fused C calls (ops.NrmsStepFn)
            params, cfg = fused
            scores, loss = ops.NrmsStepFn.apply(*params, batch["x_hist"]["title"], batch["x_cand"]["title"], batch["batch_hist"],
                                                batch["batch_cand"], batch["labels"], (B, Hmax, Cmax) + cfg)
        else:
            scores = self._forward_with_layout(batch, layout)
            loss = self._loss(scores, batch["labels"], off_c)
        cand_news_size = (off_c[1:] - off_c[:-1]).long()
        hist_news_size = (off_h[1:] - off_h[:-1]).long()
        # scores[mask_cand] (abstract_recommender.py:126-130) in ragged order, without the host sync of boolean indexing
        preds = ops.dense_to_ragged(scores, off_c, int(batch["labels"].numel()))
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
        self._val_loss_sum = getattr(self, "_val_loss_sum", 0.0) + loss.detach()
        self._val_loss_n = getattr(self, "_val_loss_n", 0) + 1
        self.val_step_outputs = self._collect_step_outputs(self.val_step_outputs, locals())
        return loss

    def on_validation_epoch_end(self):
        m = self._epoch_metrics(self.val_step_outputs, "val/")
        if getattr(self, "_val_loss_n", 0):
            # nrms_module.py:412-424: the best (lowest) epoch-mean validation loss so far, logged as a plain value
            epoch_loss = self._val_loss_sum / self._val_loss_n
            best = getattr(self, "_val_loss_best", None)
            self._val_loss_best = epoch_loss if best is None else torch.minimum(best, epoch_loss)
            self._val_loss_sum, self._val_loss_n = 0.0, 0
            self.log("val/loss_best", self._val_loss_best, sync_dist=True, prog_bar=True)
            m["val/loss_best"] = self._val_loss_best
        return m

    def test_step(self, batch: RecommendationBatch, batch_idx: int):
        (loss, preds, targets, cand_news_size, hist_news_size, target_categories, target_sentiments,
         hist_categories, hist_sentiments, user_ids, cand_news_ids) = self.model_step(batch)
        self.log("test/loss", loss, on_step=False, on_epoch=True, prog_bar=True, sync_dist=True)
        self.test_step_outputs = self._collect_step_outputs(self.test_step_outputs, locals())
        return loss

    def on_test_epoch_end(self):
        """``nrms_module.py:470-535``: ranking metrics, aspect-based diversity / personalization over categories and
        sentiments (when the test outputs carry them), and the recommendation dump of ``save_recs``."""
        out = self.test_step_outputs
        have = lambda *keys: all(k in out and len(out[k]) and torch.is_tensor(out[k][0]) for k in keys)
        extra = {}
        if have("preds", "cand_news_size", "hist_news_size"):
            preds = self._gather_step_outputs(out, "preds")
            sizes, hist_sizes = (self._gather_step_outputs(out, k) for k in ("cand_news_size", "hist_news_size"))
            for name, cand_key, hist_key, classes in (("categ", "target_categories", "hist_categories", self.num_categ_classes),
                                                      ("sent", "target_sentiments", "hist_sentiments", self.num_sent_classes)):
                if have(cand_key, hist_key):
                    m = aspect_metrics(preds, sizes, self._gather_step_outputs(out, cand_key),
                                       self._gather_step_outputs(out, hist_key), hist_sizes, classes, self.top_k_list, name)
                    extra.update({"test/" + k: v for k, v in m.items()})
        if extra:
            self.log_dict(extra, on_step=False, on_epoch=True, prog_bar=True, sync_dist=True)
        if getattr(self, "save_recs", False):
            if not have("user_ids", "cand_news_ids", "preds", "cand_news_size"):
                raise ValueError("save_recs=True needs user_ids, cand_news_ids, preds and cand_news_size among the test outputs")
            recs = self._get_recommendations(
                user_ids=self._gather_step_outputs(out, "user_ids"), news_ids=self._gather_step_outputs(out, "cand_news_ids"),
                scores=self._gather_step_outputs(out, "preds"), cand_news_size=self._gather_step_outputs(out, "cand_news_size"))
            self._save_recommendations(recommendations=recs, fpath=self.recs_fpath)
        m = self._epoch_metrics(out, "test/")
        m.update(extra)
        return m
