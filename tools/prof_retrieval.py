"""GPU box: wall-time breakdown of one retrieval step (3368 x 15913 x 2048, top-100 + CMC/mAP)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ctl_b200
from ctl_b200 import retrieval as R, synth

feats, pids, cams = synth.synth_retrieval(3368, 15913, 751, 2048, 3.0, 0)
q, g = feats[:3368].cuda(), feats[3368:].cuda()
def T(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("build_planes q+g        %.3f ms" % T(lambda: (R.build_planes(q), R.build_planes(g))))
qp, gp = R.build_planes(q), R.build_planes(g)
print("encode_identities (host) %.3f ms" % T(lambda: R.encode_identities(pids[:3368], pids[3368:], cams[:3368], cams[3368:], False)))
print("topk only (2 passes)     %.3f ms" % T(lambda: R.topk(qp, gp, 100)))
print("evaluate_streamed        %.3f ms" % T(lambda: R.evaluate_streamed(qp, gp, pids[:3368], pids[3368:], cams[:3368], cams[3368:])))
print("topk_and_eval (fused)    %.3f ms" % T(lambda: R.topk_and_eval(qp, gp, 100, pids[:3368], pids[3368:], cams[:3368], cams[3368:])))
print("dist_matrix (1 pass)     %.3f ms" % T(lambda: R.dist_matrix(q, g)))
