"""GPU box: 3 retrieval steps of one variant (argv[1]: caller | sorted_all | sorted_list) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ctl_b200  # noqa: F401
from ctl_b200 import retrieval as R, synth

variant = sys.argv[1]
NQ = 3368
feats, pids, cams = synth.synth_retrieval(NQ, 15913, 751, 2048, 3.0, 0)
q, g = feats[:NQ].cuda(), feats[NQ:].cuda()
args = (pids[:NQ], pids[NQ:], cams[:NQ], cams[NQ:])
qo, go = R.pid_order(pids[:NQ]), R.pid_order(pids[NQ:])
cache = R.PlaneCache()
if variant == "caller":
    ids = R.encode_ids(*args, False, q.device)
    step = lambda: R.topk_and_eval(R.build_planes(q), cache.get(g), 100, *args, ids=ids)
else:
    ids = R.encode_ids(*args, False, q.device, q_order=qo, g_order=go)
    tl = variant == "sorted_list"
    step = lambda: R.topk_and_eval(R.build_planes(q, order=qo), cache.get(g, order=go), 100, *args, ids=ids, tile_lists=tl)
for _ in range(3):
    step()
torch.cuda.synchronize()
print("done")
