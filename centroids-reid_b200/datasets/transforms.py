"""Device-side training transforms: the arithmetic of ReidTransforms.build_transforms(is_train=True)
(datasets/transforms/build.py:15-27) after `T.Resize`, for a whole batch in one kernel.

The reference draws its random numbers per image inside torchvision / `random` (flip: torch.rand(1) < p;
crop: torch.randint; erasing: random.uniform / random.randint with up to 100 attempts, random_erasing.py:30-55).
`sample_params` draws the same distributions from one numpy Generator (the stream of numbers differs -- the
reference's own stream depends on DataLoader worker scheduling); `augment_batch` is then a deterministic function of
(images, params) and is what the parity test pins against the oracle.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from .. import _native as N


def sample_params(batch: int, h: int, w: int, prob_flip=0.5, pad=10, re_prob=0.5, is_real=None, rng=None,
                  sl=0.02, sh=0.4, r1=0.3) -> np.ndarray:
    """int32 [batch, 8] = {flip, crop_top, crop_left, erase_row, erase_col, erase_h, erase_w, is_real}."""
    rng = rng if rng is not None else np.random.default_rng()
    out = np.zeros((batch, 8), dtype=np.int32)
    out[:, 7] = 1 if is_real is None else np.asarray(is_real, dtype=np.int32)
    for b in range(batch):
        out[b, 0] = rng.random() < prob_flip                       # T.RandomHorizontalFlip
        out[b, 1] = rng.integers(0, 2 * pad + 1)                   # T.RandomCrop on the padded image
        out[b, 2] = rng.integers(0, 2 * pad + 1)
        if rng.uniform(0, 1) >= re_prob:                           # random_erasing.py:32
            continue
        for _ in range(100):
            target_area = rng.uniform(sl, sh) * h * w
            aspect = rng.uniform(r1, 1 / r1)
            eh, ew = int(round(math.sqrt(target_area * aspect))), int(round(math.sqrt(target_area / aspect)))
            if ew < w and eh < h:
                out[b, 3] = rng.integers(0, h - eh + 1)            # random.randint is inclusive
                out[b, 4] = rng.integers(0, w - ew + 1)
                out[b, 5], out[b, 6] = eh, ew
                break
    return out


def augment_batch(images_u8: torch.Tensor, params, pixel_mean=(0.485, 0.456, 0.406), pixel_std=(0.229, 0.224, 0.225),
                  pad: int = 10) -> torch.Tensor:
    """images_u8: uint8 [B, H, W, 3] on the device (resized crops); params: int32 [B, 8] (numpy or tensor).
    Returns the normalised fp32 NCHW batch [B, 3, H, W]."""
    N.require_cuda(images_u8)
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
        raise ValueError(f"expected uint8 [B, H, W, 3], got {images_u8.dtype} {tuple(images_u8.shape)}")
    images_u8 = images_u8.contiguous()
    b, h, w, _ = images_u8.shape
    p = torch.as_tensor(np.asarray(params, dtype=np.int32) if not torch.is_tensor(params) else params)
    if tuple(p.shape) != (b, 8):
        raise ValueError(f"params must be [B, 8], got {tuple(p.shape)}")
    p = p.to(device=images_u8.device, dtype=torch.int32).contiguous()
    out = torch.empty(b, 3, h, w, dtype=torch.float32, device=images_u8.device)
    mean = (C.c_float * 3)(*[float(v) for v in pixel_mean])
    std = (C.c_float * 3)(*[float(v) for v in pixel_std])
    N.check(N.lib().ctl_augment_batch_u8(images_u8.data_ptr(), b, h, w, int(pad), p.data_ptr(), mean, std, out.data_ptr(),
                                         N.stream_ptr()))
    return out


def normalize_batch(images_u8: torch.Tensor, pixel_mean=(0.485, 0.456, 0.406), pixel_std=(0.229, 0.224, 0.225)) -> torch.Tensor:
    """The eval transform after `T.Resize` (datasets/transforms/build.py:29-33: ToTensor + Normalize) on the device:
    uint8 [B, H, W, 3] -> normalised fp32 NCHW [B, 3, H, W] (no flip / crop / erasing).  A validation loader that ships
    uint8 crops moves 4x fewer bytes over PCIe than one that normalises on the host."""
    b = images_u8.shape[0]
    key = (b, images_u8.device)
    p = _NEUTRAL.get(key)
    if p is None:
        p = torch.zeros(b, 8, dtype=torch.int32, device=images_u8.device)
        p[:, 7] = 1  # is_real
        _NEUTRAL[key] = p
    return augment_batch(images_u8, p, pixel_mean, pixel_std, pad=0)


_NEUTRAL = {}
