"""GPU box: per-phase clock64 breakdown of the conv epilogue for representative layers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ctl_b200
from ctl_b200 import _native as N

L = N.lib()
PH = ["wait_acc(MMA)", "wait_store_read", "bar1", "final_store_wait/residual", "tmem_wait", "math+sts", "fence+bar2+issue", "tile_coords"]

def run(name, n, h, w, cin, cout, k, stride, res):
    x = (torch.randn(n, h, w, cin, device="cuda") * 0.5).half()
    wt = (torch.randn(cout, k, k, cin, device="cuda") * 0.05).half()
    b = torch.zeros(cout, device="cuda")
    pad = 1 if k == 3 else 0
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    r = (torch.randn(n, ho, wo, cout, device="cuda")).half() if res else None
    out = torch.empty(n, ho, wo, cout, dtype=torch.float16, device="cuda")
    def call():
        N.check(L.ctl_conv2d_nhwc_f16(x.data_ptr(), n, h, w, cin, wt.data_ptr(), b.data_ptr(), N.ptr(r), out.data_ptr(),
                                      cout, k, stride, 1, 0, N.stream_ptr()))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    prof = torch.zeros(148 * 2 * 8, dtype=torch.int64, device="cuda")
    L.ctl_debug_set_conv_profile(prof.data_ptr())
    call()
    torch.cuda.synchronize()
    L.ctl_debug_set_conv_profile(None)
    p = prof.view(148, 2, 8).double()
    p = p[p[:, 0].sum(1) > 0]  # pair kernels fill one row per cluster
    lead, other = p[:, 0].mean(0), p[:, 1].mean(0)
    tot = lead.sum().item()
    print(f"{name}: {us:.1f} us; leader epilogue thread cycles total {tot:.0f}")
    print("   leader: " + ", ".join(f"{PH[i]}={lead[i].item():.0f}" for i in range(8)))
    print("   other : " + ", ".join(f"{PH[i]}={other[i].item():.0f}" for i in range(8)))

run("L1 conv3 64->256 +res", 256, 64, 32, 64, 256, 1, 1, True)
run("L1 down  64->256", 256, 64, 32, 64, 256, 1, 1, False)
run("L3 conv3 256->1024 +res", 256, 16, 8, 256, 1024, 1, 1, True)
run("L4 conv3 512->2048 +res", 256, 16, 8, 512, 2048, 1, 1, True)
run("L3 conv1 1024->256", 256, 16, 8, 1024, 256, 1, 1, False)
run("L3 conv2 3x3 256", 256, 16, 8, 256, 256, 3, 1, False)
run("L2 conv3 128->512 +res", 256, 32, 16, 128, 512, 1, 1, True)
run("L4 conv2 3x3 512", 256, 16, 8, 512, 512, 3, 1, False)
run("L1 conv1 256->64", 256, 64, 32, 256, 64, 1, 1, False)
