"""Drop-in for the tensor part of modelling/bases.py (ModelBase) and train_ctl_model.py
(CTLModel.training_step): same attribute names (backbone, bn, fc_query, center_loss,
contrastive_loss, xent), same hook signatures, the arithmetic on the B200 kernels.

pytorch_lightning is not a dependency of this package: the class derives from
pl.LightningModule when PL is importable and from nn.Module otherwise, so the reference's
Trainer can drive it unchanged where PL exists.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .. import _native as N
from .. import reduce as _reduce
from .. import retrieval as _R
from ..losses._fn import CTLStepFn
from ..losses.center_loss import CenterLoss
from ..losses.triplet_loss import CrossEntropyLabelSmooth, TripletLoss
from .baseline import Baseline, embed

try:  # pragma: no cover - PL is absent in the build image
    import pytorch_lightning as _pl

    _Base = _pl.LightningModule
except Exception:  # noqa: BLE001
    _Base = nn.Module

LOSS_NAMES = ("total", "query_xent", "query_triplet", "query_center", "centroid_triplet", "step_dist_ap",
              "step_dist_an", "l2_mean_centroid")


def ctl_losses(module, features, class_labels, is_real):
    """train_ctl_model.py:54-152 in one call: returns (total_loss tensor with autograd into
    features / centers / bn.weight / fc_query.weight, parts float32[8] on the device in
    LOSS_NAMES order).  No host synchronisation."""
    hp = module.hparams
    K = int(hp.DATALOADER.NUM_INSTANCE)
    B, D = features.shape
    if B % K != 0:
        raise ValueError(f"batch contract: B={B} must be P*K with K={K} (datasets/bases.py:346-406)")
    cfg = N.LossConfig(B, D, B // K, K, module.fc_query.weight.shape[0], float(hp.SOLVER.MARGIN),
                       float(hp.SOLVER.CENTER_LOSS_WEIGHT), float(hp.SOLVER.QUERY_XENT_WEIGHT),
                       float(hp.SOLVER.QUERY_CONTRASTIVE_WEIGHT), float(hp.SOLVER.CENTROID_CONTRASTIVE_WEIGHT),
                       float(module.bn.eps), float(module.bn.momentum), 0.1)
    return CTLStepFn.apply(features, module.center_loss.centers, module.bn.weight, module.fc_query.weight,
                           module.bn.bias, module.bn.running_mean, module.bn.running_var, class_labels, is_real, cfg)


class CTLModel(_Base):
    """ModelBase.__init__ (modelling/bases.py:53-90) + CTLModel (train_ctl_model.py:27-36)."""

    def __init__(self, cfg=None, test_dataloader=None, **kwargs):
        super().__init__()
        hp = dict(cfg) if cfg is not None else {}
        hp.update(kwargs)
        self.hparams_ctl = _AttrDict(hp)
        if not hasattr(self, "hparams") or not isinstance(getattr(self, "hparams", None), dict) or _Base is nn.Module:
            self.__dict__["hparams"] = self.hparams_ctl
        self.backbone = Baseline(self.hparams)
        self.contrastive_loss = TripletLoss(self.hparams.SOLVER.MARGIN, self.hparams.SOLVER.DISTANCE_FUNC)
        d_model = self.hparams.MODEL.BACKBONE_EMB_SIZE
        self.xent = CrossEntropyLabelSmooth(num_classes=self.hparams.num_classes)
        self.center_loss = CenterLoss(num_classes=self.hparams.num_classes, feat_dim=d_model,
                                      use_gpu=torch.cuda.is_available())
        self.center_loss_weight = self.hparams.SOLVER.CENTER_LOSS_WEIGHT
        self.bn = torch.nn.BatchNorm1d(d_model)
        self.bn.bias.requires_grad_(False)  # bases.py:83-84
        self.fc_query = torch.nn.Linear(d_model, self.hparams.num_classes, bias=False)
        nn.init.normal_(self.fc_query.weight, std=0.001)  # weights_init_classifier, bases.py:29-34
        self.losses_names = ["query_xent", "query_triplet", "query_center", "centroid_triplet"]
        self.losses_dict = {n: [] for n in self.losses_names}

    # -- training ---------------------------------------------------------------------------
    def training_step_from_features(self, features, class_labels, is_real):
        """Everything of training_step after `_, features = self.backbone(x)`
        (train_ctl_model.py:59) up to and including the loss assembly (:150-152)."""
        total, parts = ctl_losses(self, features, class_labels, is_real)
        return {"loss": total, "parts": parts}

    def training_step(self, batch, batch_idx, optimizer_idx=None):
        x, class_labels, camid, is_real = batch
        _, features = self.backbone(x)  # train mode: B200 training engine (differentiable w.r.t. the trunk parameters)
        return self.training_step_from_features(features, class_labels, is_real)

    def configure_optimizers(self):
        """modelling/bases.py:97-100 with the fused optimizers of ctl_b200.solver.build."""
        from ..solver.build import build_optimizer, build_scheduler

        optimizers_list = build_optimizer(self.named_parameters(), self.hparams)
        self.lr_scheduler = build_scheduler(optimizers_list[0], self.hparams)
        return optimizers_list, self.lr_scheduler

    def optimizer_step_manual(self, opt, opt_center, epoch: int = 0):
        """The tail of train_ctl_model.py:154-159 after `manual_backward`: warm-up LR rule (bases.py:115-121),
        `opt.step()`, center gradients rescaled by 1 / CENTER_LOSS_WEIGHT, `opt_center.step()`; the packed eval
        weights are invalidated because the parameters changed."""
        from ..solver.build import apply_warmup_lr

        apply_warmup_lr(opt, epoch, self.hparams)
        opt.step()
        for param in self.center_loss.parameters():
            param.grad.data *= 1.0 / self.hparams.SOLVER.CENTER_LOSS_WEIGHT
        opt_center.step()
        self.backbone.invalidate()

    # -- evaluation -------------------------------------------------------------------------
    def validation_step(self, batch, batch_idx):
        """modelling/bases.py:169-177."""
        self.backbone.eval()
        self.bn.eval()
        x, class_labels, camid, idx = batch
        with torch.no_grad():
            emb = embed(self, x)
        return {"emb": emb, "labels": class_labels, "camid": camid, "idx": idx}

    test_step = validation_step

    def validation_create_centroids(self, embeddings, labels, camids, respect_camids=False):
        """modelling/bases.py:179-262 (features stay on the device)."""
        return _reduce.validation_create_centroids(embeddings, labels, camids, self.hparams.num_query, respect_camids)

    @staticmethod
    def _calculate_centroids(vecs, dim=1):
        return _reduce._calculate_centroids(vecs, dim)

    def get_val_metrics(self, embeddings, labels, camids):
        """modelling/bases.py:264-297 without the loggers: returns (cmc, mAP, all_topk)."""
        hp = self.hparams
        respect = bool(hp.MODEL.KEEP_CAMID_CENTROIDS and hp.MODEL.USE_CENTROIDS)
        nq = hp.num_query
        emb = torch.as_tensor(embeddings).float()
        emb = emb if emb.is_cuda else emb.cuda(non_blocking=True)
        qp = _R.build_planes(emb[:nq], hp.SOLVER.DISTANCE_FUNC, hp.TEST.FEAT_NORM)
        gp = _R.build_planes(emb[nq:], hp.SOLVER.DISTANCE_FUNC, hp.TEST.FEAT_NORM)
        res = _R.evaluate_streamed(qp, gp, np.asarray(labels[:nq]), np.asarray(labels[nq:]), camids[:nq], camids[nq:],
                                   50, respect)
        for top_k, kk in zip(res.all_topk, [1, 5, 10, 20, 50]):
            print("top-k, Rank-{:<3}:{:.1%}".format(kk, top_k))
        print(f"mAP: {res.mAP}")
        return res.cmc, res.mAP, res.all_topk

    def validation_epoch_end(self, outputs):
        """modelling/bases.py:299-318 (rank-0 gating and loggers are the Trainer's business)."""
        embeddings = torch.cat([x["emb"] for x in outputs]).detach()
        labels = torch.cat([x["labels"] for x in outputs]).detach().cpu().numpy()
        camids = torch.cat([x["camid"] for x in outputs]).detach().cpu().numpy()
        if self.hparams.MODEL.USE_CENTROIDS:
            print("Evaluation is done using centroids")
            embeddings, labels, camids = self.validation_create_centroids(
                embeddings, labels, camids, respect_camids=self.hparams.MODEL.KEEP_CAMID_CENTROIDS)
        return self.get_val_metrics(embeddings, labels, camids)

    test_epoch_end = validation_epoch_end


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v
