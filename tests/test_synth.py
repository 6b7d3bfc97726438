"""CPU: the product-side synthetic generators are bit-identical to the oracle's copies (which
produced the golden vectors)."""
import numpy as np
import torch

import ctl_b200  # noqa: F401
from ctl_b200 import synth as S
from oracle import ctl_oracle as O


def test_generators_identical():
    for ibn in (False, True):
        a, b = S.make_trunk_state(3, ibn), O.make_trunk_state(3, ibn)
        assert list(a) == list(b)
        assert all(torch.equal(a[k], b[k]) for k in a)
    for dy in (False, True):
        fa, pa, ca = S.synth_retrieval(20, 90, 7, 256, 3.0, 4, dyadic=dy)
        fb, pb, cb = O.synth_retrieval(20, 90, 7, 256, 3.0, 4, dyadic=dy)
        assert torch.equal(fa, fb) and np.array_equal(pa, pb) and np.array_equal(ca, cb)
    xa, xb = S.synth_batch(6, 4, 128, 50, 2, 0.3), O.synth_batch(6, 4, 128, 50, 2, 0.3)
    assert all(torch.equal(u, v) for u, v in zip(xa, xb))
