"""Drop-in for modelling/baseline.py: Baseline(cfg).forward(x) -> (base_out, global_feat).

Parameters live in reference-layout modules (`self.base.*`); eval-mode forward runs the B200
engine, whose packed operands are rebuilt lazily whenever the parameters change
(`invalidate()`; call it after `opt.step()` / `load_state_dict`).  Train-mode forward (ResNet-50) runs the
training engine (batch-statistics BatchNorm, running statistics updated in place) and is differentiable:
`global_feat.backward()` fills `.grad` of every trunk parameter through the B200 backward kernels.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from .backbones.engine import TrunkEngine
from .backbones.engine_train import TrunkTrainer
from .backbones.resnet import ResNetParams

_LAYERS = {"resnet50": ((3, 4, 6, 3), False), "resnet101": ((3, 4, 23, 3), False), "resnet152": ((3, 8, 36, 3), False),
           "resnet50_ibn_a": ((3, 4, 6, 3), True), "resnet101_ibn_a": ((3, 4, 23, 3), True)}


class _TrunkTrainFn(torch.autograd.Function):
    """global_feat = trunk(x; parameters) with the B200 training engine; backward returns the parameter gradients
    (the input crops get no gradient, like the reference's data tensors)."""

    @staticmethod
    def forward(ctx, x, trainer, names, buffers, *tensors):
        params = dict(zip(names, tensors))
        params.update(buffers)
        ctx.trainer, ctx.names = trainer, names
        return trainer.forward(x, params)

    @staticmethod
    def backward(ctx, dfeat):
        grads = ctx.trainer.backward(dfeat)
        return (None, None, None, None) + tuple(grads.get(n) for n in ctx.names)


class Baseline(nn.Module):
    in_planes = 2048

    def __init__(self, cfg):
        super().__init__()
        name = cfg.MODEL.NAME
        if name not in _LAYERS:
            raise NotImplementedError(f"MODEL.NAME={name!r}: the B200 trunk covers the bottleneck ResNets {sorted(_LAYERS)}")
        layers, ibn = _LAYERS[name]
        self.model_name = name
        self.use_mixed_precision = cfg.USE_MIXED_PRECISION
        self.base = ResNetParams(cfg.MODEL.LAST_STRIDE, layers, ibn)
        if cfg.MODEL.PRETRAINED and not cfg.MODEL.RESUME_TRAINING and not cfg.TEST.ONLY_TEST:
            self.base.load_param(cfg.MODEL.PRETRAIN_PATH)  # modelling/baseline.py:84-87
            print("Loading pretrained ImageNet model......")
        self.gap = nn.AdaptiveAvgPool2d(1)
        self._engine = None
        self._engine_key = None
        self._trainer = None

    def invalidate(self):
        self._engine = None

    def engine(self, bn_head=None) -> TrunkEngine:
        dev = next(self.base.parameters()).device
        key = (str(dev), id(bn_head))
        if self._engine is None or self._engine_key != key:
            sd = {k: v for k, v in self.base.state_dict().items()}
            head = None
            if bn_head is not None:
                head = dict(weight=bn_head.weight, bias=bn_head.bias, running_mean=bn_head.running_mean,
                            running_var=bn_head.running_var)
            self._engine = TrunkEngine(sd, dev, ibn=self.base.ibn, last_stride=self.base.last_stride,
                                       layers=self.base.layers_cfg, bn_head=head)
            self._engine_key = key
        return self._engine

    def forward(self, x):
        """modelling/baseline.py:91-96.  base_out is returned in the reference's NCHW view."""
        if self.training:
            dev = next(self.base.parameters()).device
            if self._trainer is None or self._trainer.device != dev:
                self._trainer = TrunkTrainer(dev, last_stride=self.base.last_stride, layers=self.base.layers_cfg,
                                             graphs=os.environ.get("CTL_TRAIN_GRAPHS", "1") == "1", ibn=self.base.ibn)
            names = [k for k, _ in self.base.named_parameters()]
            tensors = [v for _, v in self.base.named_parameters()]
            buffers = {k: v for k, v in self.base.named_buffers() if "running" in k}
            feat = _TrunkTrainFn.apply(x, self._trainer, names, buffers, *tensors)
            self.invalidate()  # running statistics changed: the eval engine must refold them
            for k, v in self.base.named_buffers():
                if k.endswith("num_batches_tracked"):
                    v += 1
            return None, feat  # callers use only global_feat (train_ctl_model.py:55); base_out is not kept
        out = self.engine().forward(x, want_base=True)
        return out["base_out_nhwc"].permute(0, 3, 1, 2), out["global_feat"]


def embed(pl_module, x):
    """ModelBase.validation_step's arithmetic (modelling/bases.py:169-177) ==
    inference_utils._inference (inference/inference_utils.py:104-113): eval trunk -> GAP ->
    eval BatchNorm1d, fused into the engine's last kernel."""
    eng = pl_module.backbone.engine(bn_head=pl_module.bn)
    return eng.forward(x, want_emb=True)["emb"]
