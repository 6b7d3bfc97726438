"""torchrun --nproc-per-node N tools/test_sharded_gpu.py : gallery-sharded retrieval over NCCL must equal
the single-GPU result (top-k under the canonical key order, CMC / mAP), and times BASELINE config 5's
shape scaled to N GPUs (50k queries x 25k*N gallery rows)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import ctl_b200
from ctl_b200 import retrieval as R, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
# ---- correctness on a small problem ---------------------------------------------------------
nq, ng, k = 128 * world, 6000, 50
feats, pids, cams = synth.synth_retrieval(nq, ng, 300, 512, 2.0, 5, num_cams=4)
q, g = feats[:nq], feats[nq:]
shard = np.array_split(np.arange(ng), world)[rank]
q_local = q[rank * (nq // world):(rank + 1) * (nq // world)].to(dev)
idx, dst = R.topk_sharded(q_local, g[shard].to(dev), k, int(shard[0]), None)
ridx, rdst = R.topk_similar(q.to(dev), g.to(dev), k)
ok1 = torch.equal(idx, ridx) and torch.equal(dst, rdst)
qp = R.build_planes(q.to(dev))
res = R.evaluate_streamed(qp, R.build_planes(g[shard].to(dev)), pids[:nq], pids[nq:][shard], cams[:nq], cams[nq:][shard],
                          50, False, g_index_offset=int(shard[0]), group=dist.group.WORLD, total_gallery=ng)
ref = R.evaluate_streamed(qp, R.build_planes(g.to(dev)), pids[:nq], pids[nq:], cams[:nq], cams[nq:])
ok2 = np.array_equal(res.cmc, ref.cmc) and res.mAP == ref.mAP and np.array_equal(res.ranks[:, :ref.ranks.shape[1]], ref.ranks) \
    if res.ranks.shape[1] >= ref.ranks.shape[1] else False
flags = torch.tensor([int(ok1), int(ok2)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded top-k == single GPU: {bool(flags[0])};  sharded CMC/mAP/ranks == single GPU: {bool(flags[1])}")
# ---- config-5 shape, weak-scaled: 50k queries, 25k gallery rows per GPU ------------------------------
Q, GL, D, K = 50_000, 25_000, 2048, 100
gen = torch.Generator(device=dev).manual_seed(100 + rank)
ql = torch.nn.functional.normalize(torch.randn(Q // world, D, device=dev, generator=gen), dim=1)
gl = torch.nn.functional.normalize(torch.randn(GL, D, device=dev, generator=gen), dim=1)
for _ in range(2):
    R.topk_sharded(ql, gl, K, rank * GL, None)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
n = 3
for _ in range(n):
    out = R.topk_sharded(ql, gl, K, rank * GL, None)
torch.cuda.synchronize(); dist.barrier()
dt = (time.perf_counter() - t0) / n
if rank == 0:
    print(f"top-{K} of {Q} x {GL * world} x {D} on {world} GPU(s): {dt*1e3:.1f} ms -> {Q*GL*world/dt/1e9:.1f} Gpairs/s")
dist.destroy_process_group()
