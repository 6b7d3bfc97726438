"""Synthetic workloads of the BASELINE configs (SURVEY.md section 8d 'Synthetic inputs'):
random-init trunk weights in the reference's state_dict layout, N(0,1) crops, clustered unit
vectors for retrieval, pid-major CTL batches.  No datasets or checkpoints exist offline.

(tests/test_synth.py asserts these generators are bit-identical to the oracle's own copies,
which is what the golden vectors were generated from.)
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

R50_LAYERS = (3, 4, 6, 3)


def make_trunk_state(seed=0, ibn=False, layers=R50_LAYERS, randomize_bn=True):
    """Reference-layout trunk weights: conv ~ N(0, sqrt(2/(k*k*Cout))) (resnet.py:156-164),
    BN affine / running statistics randomised so folding is exercised; bn3 gains are small, as
    in trained nets, so the residual sum stays O(1) and inside fp16 range."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def conv(name, cout, cin, k):
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (k * k * cout))

    def bn(name, c):
        if randomize_bn:
            gain = 0.25 if name.endswith("bn3") else 1.0
            sd[name + ".weight"] = gain * (0.5 + torch.rand(c, generator=g))
            sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
            sd[name + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
            sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        else:
            sd[name + ".weight"] = torch.ones(c)
            sd[name + ".bias"] = torch.zeros(c)
            sd[name + ".running_mean"] = torch.zeros(c)
            sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def inorm(name, c):
        sd[name + ".weight"] = 0.5 + torch.rand(c, generator=g) if randomize_bn else torch.ones(c)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g) if randomize_bn else torch.zeros(c)

    conv("conv1", 64, 3, 7)
    bn("bn1", 64)
    inplanes = 64
    for li, (planes, nblk) in enumerate(zip((64, 128, 256, 512), layers), start=1):
        for b in range(nblk):
            p = f"layer{li}.{b}"
            conv(p + ".conv1", planes, inplanes, 1)
            if ibn and planes != 512:
                half = planes // 2
                inorm(p + ".bn1.IN", half)
                bn(p + ".bn1.BN", planes - half)
            else:
                bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3)
            bn(p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, 1)
            bn(p + ".bn3", planes * 4)
            if b == 0:
                conv(p + ".downsample.0", planes * 4, inplanes, 1)
                bn(p + ".downsample.1", planes * 4)
                inplanes = planes * 4
    if ibn:
        sd["fc.weight"] = torch.zeros(1000, 2048)
        sd["fc.bias"] = torch.zeros(1000)
    return sd


def make_head_bn(seed=0, dim=2048):
    """ModelBase.bn (BatchNorm1d(2048), modelling/bases.py:83) with randomised eval statistics."""
    g = torch.Generator().manual_seed(10_000 + seed)
    return dict(weight=0.5 + torch.rand(dim, generator=g), bias=torch.zeros(dim),
                running_mean=0.1 * torch.randn(dim, generator=g), running_var=0.5 + torch.rand(dim, generator=g))


def synth_retrieval(num_q, num_g, num_ids, dim=2048, sigma=3.0, seed=0, num_cams=6, dyadic=False):
    """Clustered unit vectors normalize(c_pid + sigma N(0,I)); dyadic=True: a coarse dyadic grid
    on which every dot product is exact in fp32 and in the fp16-split tensor-core arithmetic."""
    g = torch.Generator().manual_seed(seed)
    n = num_q + num_g
    pids = torch.randint(0, num_ids, (n,), generator=g)
    cams = torch.randint(0, num_cams, (n,), generator=g)
    if dyadic:
        centres = torch.randint(-2, 3, (num_ids, dim), generator=g).float()
        noise = torch.randint(-2, 3, (n, dim), generator=g).float()
        keep = (torch.rand(n, dim, generator=g) < 0.5).float()
        feats = (centres[pids] * keep + noise * (1 - keep)) / 16.0
    else:
        centres = torch.randn(num_ids, dim, generator=g)
        feats = centres[pids] + sigma * torch.randn(n, dim, generator=g)
        feats = F.normalize(feats, dim=1)
    return feats, pids.numpy().astype(np.int64), cams.numpy().astype(np.int64)


def synth_batch(P, K, dim=2048, num_classes=751, seed=0, pad_fraction=0.0, scale=1.0, pid_offset=0.15):
    """Post-trunk inputs of one CTL step under the batch contract A0 (pid-major blocks of K,
    padded rows at the end of a block, >= 2 real rows per pid)."""
    g = torch.Generator().manual_seed(seed)
    B = P * K
    pid_pool = torch.randperm(num_classes, generator=g)[:P]
    labels = pid_pool.repeat_interleave(K)
    feats = scale * torch.randn(B, dim, generator=g)
    feats = feats + scale * pid_offset * torch.randn(P, dim, generator=g).repeat_interleave(K, 0)
    is_real = torch.ones(B, dtype=torch.bool)
    if pad_fraction > 0:
        npad = max(1, int(round(P * pad_fraction)))
        for c in torch.randperm(P, generator=g)[:npad].tolist():
            drop = int(torch.randint(1, max(2, min(3, K - 1)), (1,), generator=g))
            drop = min(drop, K - 2)
            if drop > 0:
                is_real[c * K + K - drop: (c + 1) * K] = False
    return feats, labels, is_real
