"""GPU: training-side trunk kernels (weight gradient, batch-statistics BatchNorm forward/backward, ...)
against float64 torch autograd of the same fp16-rounded operands."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

WGRAD_SHAPES = [
    # n, h, w, cin, cout, k, stride
    (2, 16, 8, 64, 64, 1, 1),
    (3, 16, 8, 64, 64, 3, 1),
    (2, 16, 8, 256, 128, 1, 1),
    (2, 16, 16, 128, 128, 3, 2),
    (2, 12, 20, 64, 256, 1, 2),
    (5, 8, 4, 512, 192, 3, 1),
    (4, 32, 16, 128, 512, 1, 1),
]


@pytest.mark.parametrize("shape", WGRAD_SHAPES)
def test_conv_wgrad(shape):
    from ctl_b200 import _native as N

    L = N.lib()
    n, h, w, cin, cout, k, stride = shape
    g = torch.Generator().manual_seed(hash(shape) % 1000)
    pad = 1 if k == 3 else 0
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    x = (torch.randn(n, h, w, cin, generator=g)).half()
    dy = (torch.randn(n, ho, wo, cout, generator=g) * 0.5).half()
    ref = torch.nn.grad.conv2d_weight(x.double().permute(0, 3, 1, 2), (cout, cin, k, k), dy.double().permute(0, 3, 1, 2),
                                      stride=stride, padding=pad).permute(0, 2, 3, 1)  # [cout][k][k][cin]
    xd, dyd = x.cuda(), dy.cuda()
    nbytes = L.ctl_conv2d_wgrad_workspace_bytes(n, h, w, cin, cout, k, stride)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dw = torch.full((cout, k, k, cin), float("nan"), device="cuda")
    N.check(L.ctl_conv2d_wgrad_nhwc_f16(xd.data_ptr(), n, h, w, cin, dyd.data_ptr(), cout, k, stride, ws.data_ptr(),
                                        nbytes, dw.data_ptr(), N.stream_ptr()))
    torch.cuda.synchronize()
    got = dw.cpu().double()
    assert torch.isfinite(got).all()
    # fp32 accumulation of exact fp16 products: error ~ sqrt(K) * 2^-24 * |terms|
    assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + 1e-4
    # deterministic: a second call reproduces the bits
    dw2 = torch.empty_like(dw)
    N.check(L.ctl_conv2d_wgrad_nhwc_f16(xd.data_ptr(), n, h, w, cin, dyd.data_ptr(), cout, k, stride, ws.data_ptr(),
                                        nbytes, dw2.data_ptr(), N.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(dw, dw2)
