"""Drop-in for losses/center_loss.py of the reference."""
from __future__ import annotations

import torch
import torch.nn as nn

from ._fn import CenterLossFn, raise_if_poisoned


class CenterLoss(nn.Module):
    """losses/center_loss.py:4-45.  Center loss (Wen et al., ECCV 2016) with the reference's
    quirk kept: the [B, C] masked matrix is clamped element-wise, so every masked zero adds
    1e-12 to the sum.  The kernel gathers the B needed center rows instead of forming the
    [B, C] GEMM (2 MB instead of 6.2 MB of traffic at B=256, C=751)."""

    def __init__(self, num_classes=751, feat_dim=2048, use_gpu=True):
        super().__init__()
        self.num_classes = num_classes
        self.feat_dim = feat_dim
        self.use_gpu = use_gpu
        centers = torch.randn(self.num_classes, self.feat_dim)
        if self.use_gpu:
            centers = centers.cuda()
        self.centers = nn.Parameter(centers)

    def forward(self, x, labels):
        assert x.size(0) == labels.size(0), "features.size(0) is not equal to labels.size(0)"
        loss = CenterLossFn.apply(x, self.centers, labels)
        # a label outside [0, num_classes) would index `centers` out of bounds: the kernel clamps it and reports it as a
        # NaN with a payload; this stand-alone class reads the loss back (one float) and raises.  The fused training step
        # (modelling/ctl_model.py) does not go through here.
        raise_if_poisoned(loss, "CenterLoss")
        return loss
