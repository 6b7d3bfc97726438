"""GPU box: a few top-k calls at the Market1501 shape (for ncu captures of dist_gemm_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ctl_b200
from ctl_b200 import retrieval as R, synth
feats, pids, cams = synth.synth_retrieval(3368, 15913, 751, 2048, 3.0, 0)
q, g = feats[:3368].cuda(), feats[3368:].cuda()
qp, gp = R.build_planes(q), R.build_planes(g)
for _ in range(3):
    R.topk_and_eval(qp, gp, 100, pids[:3368], pids[3368:], cams[:3368], cams[3368:])
torch.cuda.synchronize()
print("done")
