"""GPU: device-side training transforms (flip / pad-crop / normalise / random erasing) vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import ctl_oracle as O

pytestmark = pytest.mark.gpu


def test_augment_batch_matches_oracle_and_reference_random_erasing():
    from ctl_b200.datasets import transforms as T

    rng = np.random.default_rng(3)
    B, H, W = 9, 32, 20
    imgs = torch.from_numpy(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8))
    is_real = np.ones(B, dtype=np.int32)
    is_real[4] = 0
    params = T.sample_params(B, H, W, prob_flip=0.5, pad=4, re_prob=0.7, is_real=is_real, rng=rng)
    params[0] = [1, 0, 8, 0, 0, H - 1, W - 1, 1]   # extreme crop corner + the largest erasable rectangle
    params[1] = [0, 8, 0, 5, 3, 0, 0, 1]           # erase_h == 0: no erasing
    assert params[:, 1:3].min() >= 0 and params[:, 1:3].max() <= 8 and (params[:, 5] < H).all() and (params[:, 6] < W).all()
    got = T.augment_batch(imgs.cuda(), params, pad=4).cpu()
    ref = O.augment_batch(imgs, params, pad=4)
    assert torch.equal(got[4], torch.zeros(3, H, W))
    assert torch.equal(got, ref)  # same IEEE operations in the same order: bit-identical
    # statistics of the sampler: flip about half, erase about re_prob, crops cover the whole offset range
    big = T.sample_params(4000, 256, 128, rng=np.random.default_rng(0))
    assert abs(big[:, 0].mean() - 0.5) < 0.03 and abs((big[:, 5] > 0).mean() - 0.5) < 0.03
    assert big[:, 1].min() == 0 and big[:, 1].max() == 20 and big[:, 2].min() == 0 and big[:, 2].max() == 20
    area = (big[:, 5] * big[:, 6])[big[:, 5] > 0] / (256 * 128)
    assert area.min() >= 0.015 and area.max() <= 0.41
    with pytest.raises(ValueError):
        T.augment_batch(imgs.cuda().float(), params)
