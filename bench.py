#!/usr/bin/env python
"""Benchmark of the B200-native centroid-triplet re-ID hot path (driver contract: ONE JSON line).

    python bench.py --gpus N --steps K --warmup W                      (our CUDA path)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                               (the reference's CPU algorithm)

Metric (BASELINE.json): embeddings/sec @256x128 -- one "step" = one pass of the eval embedding
path (trunk -> global average pool -> BatchNorm1d, modelling/bases.py:169-177) over one batch
of 256 synthetic 256x128 crops per GPU, fp16 activations / fp32 accumulation, random-init
ResNet50 weights of the reference architecture.  At N > 1 every rank embeds its own batch and the
per-rank embeddings are all-gathered once per step over NCCL (weak scaling).  The second
BASELINE metric, Q x G top-k pairs/sec (config 3: 3368 x 15913 x 2048, top-100 + CMC/mAP), is
reported in the same line under "retrieval" (and is the primary metric with --workload retrieval).

Only the `cpu_baseline` leg and `--impl reference` execute anything under oracle/ (the CPU
restatement of the reference, timed on the host cores as the baseline).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

BATCH = 256
H, W = 256, 128
GFLOP_PER_IMG = 8.1065  # sum over the 53 convolutions, 256x128, last_stride 1
RET_Q, RET_G, RET_D, RET_K, RET_IDS = 3368, 15913, 2048, 100, 751


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md).  The sampler
    process is started once (nvidia-smi needs ~0.5 s to come up) and polls every 20 ms; `window()`
    marks the wall-clock intervals of the timed loops and only samples inside them are summarised."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc, self.windows = gpu_index, [], None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.7)
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self):
        sampler = self

        class _W:
            def __enter__(self):
                self.t0 = time.time()

            def __exit__(self, *exc):
                sampler.windows.append((self.t0, time.time()))
                return False

        return _W()

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        for t, r in self.rows:
            if self.windows and not any(a - 0.01 <= t <= b + 0.03 for a, b in self.windows):
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no nvidia-smi sample inside the timed windows"],
                    "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def timed_steps(step_fn, steps, warmup, world):
    """W warm-ups, then EXACTLY `steps` steps between barrier + synchronize; device time via CUDA
    events on the launching stream, max over ranks."""
    import torch.distributed as dist

    for i in range(warmup):
        step_fn(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step_fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


# ----------------------------------------------------------------------------------------------
# embedding workload (metric M1)
# ----------------------------------------------------------------------------------------------

def build_engine(device):
    import ctl_b200  # noqa: F401
    from ctl_b200 import synth
    from ctl_b200.modelling.backbones.engine import TrunkEngine

    return TrunkEngine(synth.make_trunk_state(seed=0), device, ibn=False, last_stride=1, bn_head=synth.make_head_bn(0))


def run_embed(args, world, rank, local):
    import torch.distributed as dist

    dev = torch.device("cuda", local)
    eng = build_engine(dev)
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    n_rot = 4  # 4 x 100.7 MB of inputs > 126 MB L2; activations (hundreds of MB per layer) never fit anyway
    host = [torch.randn(BATCH, 3, H, W, generator=gen).pin_memory() for _ in range(n_rot)]
    dev_in = [h.to(dev) for h in host]
    gathered = [torch.empty(BATCH, 2048, device=dev) for _ in range(world)] if world > 1 else None
    from ctl_b200.modelling.backbones.engine import GraphedForward

    graphs = [GraphedForward(eng, d, want_emb=True) for d in dev_in]  # one CUDA graph per rotating input

    def step(i):
        emb = graphs[i % n_rot]()["emb"]
        if world > 1:
            dist.all_gather(gathered, emb)
        return emb

    clk = ClockSampler(local)
    clk.__enter__()
    for i in range(args.warmup):
        step(i)
    with clk.window():
        ms = timed_steps(step, args.steps, 0, world)
    launches = graphs[0].launches * args.steps
    value = world * BATCH * args.steps / (ms / 1e3)

    # ---- end to end: pinned host crops -> H2D -> forward -> D2H embeddings, double-buffered ----
    copy_stream = torch.cuda.Stream(device=dev)
    out_host = [torch.empty(BATCH, 2048).pin_memory() for _ in range(2)]
    stage = [torch.empty(BATCH, 3, H, W, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    stage_graphs = [GraphedForward(eng, st, want_emb=True) for st in stage]

    def prefetch(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[b])  # the forward that last read this staging buffer
            stage[b].copy_(host[i % n_rot], non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_loop(n_steps, first):
        prefetch(first)
        for j in range(n_steps):
            i = first + j
            b = i % 2
            if j + 1 < n_steps:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[b])
            emb = stage_graphs[b]()["emb"]
            done[b].record()
            if world > 1:
                dist.all_gather(gathered, emb)
            out_host[b].copy_(emb, non_blocking=True)

    for b in range(2):
        done[b].record()
    e2e_loop(args.warmup, 0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    with clk.window():
        e2e_loop(args.steps, args.warmup)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = time.perf_counter() - t0
    clk.__exit__(None, None, None)
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * BATCH * args.steps / e2e_s, "unit": "embeddings/s",
           "h2d_bytes_per_step": BATCH * 3 * H * W * 4, "d2h_bytes_per_step": BATCH * 2048 * 4,
           "note": "public API TrunkEngine.forward on pinned host crops; H2D of step i+1 overlaps compute of step i"}

    # ---- roofline of the dominant kernel (conv_gemm): per-launch CUDA events, one profiled pass ----
    roof = None
    if rank == 0:
        eng.profile = []
        for i in range(3):
            eng.profile.clear()
            eng.forward(dev_in[i % n_rot], want_emb=True)
        torch.cuda.synchronize()
        agg = {}
        for name, fl, by, a, b in eng.profile:
            d = agg.setdefault(name, [0.0, 0.0, 0.0, 0])
            d[0] += fl
            d[1] += by
            d[2] += a.elapsed_time(b)
            d[3] += 1
        eng.profile = None
        pk = peaks()
        tot_ms = sum(v[2] for v in agg.values())
        c = agg["conv_gemm"]
        ach = c[0] / (c[2] * 1e-3) / 1e12
        roof = {"kernel": "conv_gemm kernels (52 launches/step, fused conv+BN+residual+ReLU)", "bound": "tensor",
                "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                "peak_source": pk["src"] + ", bf16 sustained", "traffic": conv_traffic(),
                "share_of_step": c[2] / tot_ms,
                "hbm_achieved_gbs": c[1] / (c[2] * 1e-3) / 1e9, "hbm_peak_gbs": pk["hbm"],
                "other_kernels_ms": {k: round(v[2], 4) for k, v in agg.items() if k != "conv_gemm"},
                "conv_ms": round(c[2], 4)}
    return ms, value, launches, e2e, roof, clk.summary()


def conv_traffic():
    """DRAM bytes (read + write) of the 52 conv launches of one bs-256 forward, from the committed ncu metrics
    pass (profiles/conv_traffic.json, written by tools/ncu_traffic.py on the GPU box); None if absent."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "conv_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)["dram_bytes_per_step"]
    except (OSError, KeyError, ValueError):
        return None


def run_train_step(local, steps=5, warmup=2, model_name="resnet50", size=(256, 128), P=16, K=16, world=1):
    """BASELINE config 2 (and, with model_name="resnet50_ibn_a", size=(320, 320), P=32, K=4, world=8, config 4):
    one complete CTL training iteration per step -- train-mode trunk forward (batch-stat BN) -> fused
    CTL/center/xent/triplet loss step -> backward through the loss and the trunk (all parameter gradients) ->
    [world > 1: NCCL mean all-reduce of the gradients in flat buckets] -> fused Adam + center-SGD step.
    Every rank trains on its own P x K batch (weak scaling, like the reference's DDP).  Device-timed with CUDA
    events; the caller takes the max over ranks."""
    import ctl_b200  # noqa: F401
    from ctl_b200.modelling.ctl_model import CTLModel

    dev = torch.device("cuda", local)

    class _C(dict):
        __getattr__ = dict.__getitem__

    if True:
        cfg = _C(MODEL=_C(NAME="resnet50", LAST_STRIDE=1, PRETRAINED=False, PRETRAIN_PATH="", BACKBONE_EMB_SIZE=2048,
                          USE_CENTROIDS=False, KEEP_CAMID_CENTROIDS=True, RESUME_TRAINING=False),
                 SOLVER=_C(MARGIN=0.5, DISTANCE_FUNC="euclidean", CENTER_LOSS_WEIGHT=5e-4, QUERY_XENT_WEIGHT=1.0,
                           QUERY_CONTRASTIVE_WEIGHT=1.0, CENTROID_CONTRASTIVE_WEIGHT=1.0, OPTIMIZER_NAME="Adam",
                           BASE_LR=1e-4, WEIGHT_DECAY=5e-4, CENTER_LR=0.5, LR_SCHEDULER_NAME="multistep_lr",
                           LR_STEPS=(40, 70), GAMMA=0.1, USE_WARMUP_LR=True, WARMUP_EPOCHS=10),
                 DATALOADER=_C(NUM_INSTANCE=K), TEST=_C(FEAT_NORM=True, ONLY_TEST=False, VISUALIZE="no"),
                 USE_MIXED_PRECISION=True)
    torch.manual_seed(0)
    from ctl_b200 import parallel

    model = CTLModel(cfg, num_classes=751, num_query=0).to(dev).train()
    g = torch.Generator().manual_seed(1234 + local)
    x = torch.randn(P * K, 3, size[0], size[1], generator=g).to(dev)
    labels = torch.arange(P).repeat_interleave(K).to(dev)
    cam = torch.zeros(P * K, dtype=torch.long, device=dev)
    is_real = torch.ones(P * K, dtype=torch.bool, device=dev)

    (opt, opt_center), _ = model.configure_optimizers()

    def step():
        for p_ in model.parameters():
            p_.grad = None
        out = model.training_step((x, labels, cam, is_real), 0)
        out["loss"].backward()
        if world > 1:
            parallel.allreduce_gradients(model.parameters())  # one mean all-reduce over NCCL, flat fp32 buckets
        model.optimizer_step_manual(opt, opt_center, epoch=0)  # fused Adam + center SGD (solver/build.py)
        return out["loss"]

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    gflop = GFLOP_PER_IMG if (model_name == "resnet50" and tuple(size) == (256, 128)) else None
    return {"metric": f"CTL training step images/sec ({model_name} {size[0]}x{size[1]}, {P} ids x {K} instances per GPU, "
                      "fwd+loss+bwd+optimizer)",
            "value": world * P * K / ms * 1e3, "unit": "images/s", "ms_per_step": ms, "steps": steps, "loss": float(loss.detach()),
            "tflops": (3 * world * P * K * gflop / ms) if gflop else None,
            "note": "3 x forward FLOPs per image; includes the gradient all-reduce (N > 1) and the fused Adam / center-SGD step",
            "peak_mem_gib": torch.cuda.max_memory_allocated(dev) / 2 ** 30}


# ----------------------------------------------------------------------------------------------
# retrieval workload (metric M2)
# ----------------------------------------------------------------------------------------------

def run_retrieval(args, world, rank, local, steps=None, warmup=None):
    """Config 3 on ONE GPU (3368 x 15913 x 2048, top-100 + CMC/mAP).  With world > 1 the gallery is
    sharded: queries all-gathered, per-rank top-k merged (retrieval.topk_sharded)."""
    import ctl_b200  # noqa: F401
    from ctl_b200 import retrieval as R
    from ctl_b200 import synth

    steps = steps or args.steps
    warmup = warmup or args.warmup
    dev = torch.device("cuda", local)
    feats, pids, cams = synth.synth_retrieval(RET_Q, RET_G, RET_IDS, RET_D, 3.0, 0)
    qh, gh = feats[:RET_Q].contiguous().pin_memory(), feats[RET_Q:].contiguous().pin_memory()
    q, g = qh.to(dev), gh.to(dev)
    box = {}
    # like the features, the identity arrays of the validation set are resident on the device for `value`
    # (they are re-encoded from the host arrays every step in the e2e loop below)
    ids = R.encode_ids(pids[:RET_Q], pids[RET_Q:], cams[:RET_Q], cams[RET_Q:], False, dev)

    def step(i):
        qp, gp = R.build_planes(q), R.build_planes(g)
        idx, dst, res = R.topk_and_eval(qp, gp, RET_K, pids[:RET_Q], pids[RET_Q:], cams[:RET_Q], cams[RET_Q:], ids=ids)
        box["res"] = res

    # the step contains a host read-back (CMC/mAP reduction), so wall time on a quiet stream == device time
    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps

    def e2e_step(i):
        qd, gd = qh.to(dev, non_blocking=True), gh.to(dev, non_blocking=True)
        qp, gp = R.build_planes(qd), R.build_planes(gd)
        idx, dst, res = R.topk_and_eval(qp, gp, RET_K, pids[:RET_Q], pids[RET_Q:], cams[:RET_Q], cams[RET_Q:])
        return idx.cpu(), dst.cpu(), res

    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        e2e_step(i)
    torch.cuda.synchronize()
    dte = (time.perf_counter() - t0) / steps
    # GEMM kernel alone: two passes per step, 3 fp16 MMAs per pair-element
    qp, gp = R.build_planes(q), R.build_planes(g)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = torch.empty(1, device=dev)
    for _ in range(2):
        R.topk(qp, gp, RET_K)
    torch.cuda.synchronize()
    import ctypes as C
    from ctl_b200 import _native as N
    gmin = torch.empty(RET_Q, (RET_G + 15) // 16, device=dev)
    desc = N.PassDesc(gmin=gmin.data_ptr())
    e0.record()
    for _ in range(5):
        N.check(N.lib().ctl_dist_pass(qp.ptr, RET_Q, gp.ptr, RET_G, RET_D, qp.flags, C.byref(desc), N.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    pass_ms = e0.elapsed_time(e1) / 5
    pk = peaks()
    flops = 2.0 * RET_Q * RET_G * RET_D  # algorithmic (2*D flop per pair); the kernel issues 3 fp16 products
    ach = flops / (pass_ms * 1e-3) / 1e12
    return {
        "metric": "QxG top-k pairs/sec (3368x15913x2048, top-100 + CMC/mAP)", "value": RET_Q * RET_G / dt,
        "unit": "pairs/s", "ms_per_step": dt * 1e3, "mAP": box["res"].mAP, "rank1": float(box["res"].cmc[0]),
        "e2e": {"value": RET_Q * RET_G / dte, "unit": "pairs/s", "h2d_bytes_per_step": (RET_Q + RET_G) * RET_D * 4,
                "d2h_bytes_per_step": RET_Q * RET_K * 12},
        "roofline": {"kernel": "dist_gemm_kernel (split-fp16 x3 tcgen05, one pass)", "bound": "tensor",
                     "achieved": ach, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": ach / pk["tf_burst"],
                     "peak_source": pk["src"] + ", bf16 burst", "pass_ms": pass_ms, "traffic": None,
                     "tensor_pipe_tflops": 3 * ach,
                     "note": "achieved = algorithmic 2*Q*G*D flop per pass; the fp32-equivalent split issues 3 fp16 "
                             "MMA products per element (tensor_pipe_tflops = 3 x achieved, %.2f of the burst peak)"
                             % (3 * ach / pk["tf_burst"])},
        "gpu_launches_per_step": 9,
    }


# ----------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle restatement of the reference, on the host cores
# ----------------------------------------------------------------------------------------------

def _host_threads():
    """All the host threads torch can use productively: its own default is one per physical core;
    hyper-thread oversubscription (os.cpu_count()) was measured 19x SLOWER on the 128-thread box."""
    n = max(1, (os.cpu_count() or 2) // 2)  # torchrun exports OMP_NUM_THREADS=1: set the count explicitly
    torch.set_num_threads(n)
    return n


def cpu_embed(n_images, reps):
    from oracle import ctl_oracle as O  # the one place the bench executes the oracle

    _host_threads()
    sd = O.make_trunk_state(seed=0)
    g = torch.Generator().manual_seed(10_000)
    bn = dict(weight=0.5 + torch.rand(2048, generator=g), bias=torch.zeros(2048),
              running_mean=0.1 * torch.randn(2048, generator=g), running_var=0.5 + torch.rand(2048, generator=g))
    x = torch.randn(n_images, 3, H, W, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        O.embed_forward(x[:8], sd, bn)  # warm-up
        t0 = time.perf_counter()
        for _ in range(reps):
            O.embed_forward(x, sd, bn)
        dt = time.perf_counter() - t0
    return n_images * reps / dt, dt


def cpu_retrieval(nq):
    from oracle import ctl_oracle as O

    _host_threads()
    feats, pids, cams = O.synth_retrieval(RET_Q, RET_G, RET_IDS, RET_D, 3.0, 0)
    q, g = feats[:nq], feats[RET_Q:]
    t0 = time.perf_counter()
    cmc, mAP, _ = O.r1_map_compute(torch.cat((q, g)), np.concatenate((pids[:nq], pids[RET_Q:])),
                                   np.concatenate((cams[:nq], cams[RET_Q:])), nq)
    dt = time.perf_counter() - t0
    return nq * RET_G / dt, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="embed", choices=["embed", "retrieval", "train"])
    ap.add_argument("--train-model", default="resnet50", choices=["resnet50", "resnet50_ibn_a"])
    ap.add_argument("--train-size", default="256x128", help="HxW of the training crops (config 4: 320x320)")
    ap.add_argument("--train-pk", default="16x16", help="ids x instances per GPU (config 4: 32x4)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary metric and the CPU baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world, rank, local = dist_env()

    workload_name = (f"resnet50 eval embedding forward (trunk->GAP->BN1d), {BATCH} synthetic 256x128 crops per GPU, "
                     "random-init weights")
    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's own algorithm on the host cores; each step = a bounded sample of the workload
        if args.workload == "embed":
            n = 32
            v, dt = cpu_embed(n, max(1, min(args.steps, 6)))
            line = {"metric": "embeddings/sec @256x128", "value": v, "unit": "embeddings/s", "sample": f"{n} of {BATCH} crops per step"}
        else:
            v, dt = cpu_retrieval(256)
            line = {"metric": "QxG top-k pairs/sec", "value": v, "unit": "pairs/s", "sample": f"256 of {RET_Q} queries x {RET_G} gallery"}
        line.update({"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                     "ms_per_step": dt * 1e3 / max(1, min(args.steps, 6)), "higher_is_better": True, "scaling": "weak",
                     "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name},
                     "cpu_baseline": {"value": v, "unit": line["unit"], "cores": _host_threads(), "kind": "port",
                                      "sample": line["sample"]},
                     "e2e": {"value": v, "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        if args.workload == "embed":
            ms, value, launches, e2e, roof, clocks = run_embed(args, world, rank, local)
            line = {"metric": "embeddings/sec @256x128", "value": value, "unit": "embeddings/s",
                    "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
                    "data": "synthetic",
                    "config": {"workload": workload_name, "batch_per_gpu": BATCH, "global_batch": BATCH * world,
                               "model": "resnet50 last_stride=1", "gflop_per_embedding": GFLOP_PER_IMG,
                               "l2": "4 rotating input batches (403 MB) > 126 MB L2; per-layer activations 67-268 MB",
                               "parallelism": f"dp{world}: per-rank batches + one NCCL all-gather of embeddings"},
                    "tflops": value * GFLOP_PER_IMG / 1e3, "roofline": roof, "e2e": e2e, "gpu_launches": launches,
                    "clocks": clocks}
            if rank == 0 and not args.no_secondary and world == 1:
                line["retrieval"] = run_retrieval(args, world, rank, local, steps=5, warmup=3)
                line["train_step"] = run_train_step(local)
                v, dt = cpu_embed(64, 2)
                line["cpu_baseline"] = {"value": v, "unit": "embeddings/s", "cores": _host_threads(), "kind": "port",
                                        "sample": f"128 crops (2 x 64) through oracle.embed_forward, torch-CPU fp32, {dt:.1f} s"}
                rv, rdt = cpu_retrieval(128)
                line["retrieval"]["cpu_baseline"] = {"value": rv, "unit": "pairs/s", "cores": _host_threads(), "kind": "port",
                                                     "sample": f"128 of {RET_Q} queries x {RET_G} gallery through oracle.r1_map_compute, {rdt:.1f} s"}
        elif args.workload == "train":
            hh, ww = (int(v) for v in args.train_size.split("x"))
            pp, kk = (int(v) for v in args.train_pk.split("x"))
            if world > 1:
                dist.barrier()
            with ClockSampler(local) as clk:
                with clk.window():
                    r = run_train_step(local, steps=args.steps, warmup=args.warmup, model_name=args.train_model,
                                       size=(hh, ww), P=pp, K=kk, world=world)
            line = {"metric": r["metric"], "value": r["value"], "unit": r["unit"], "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                    "config": {"workload": f"CTL training iteration, {args.train_model} {hh}x{ww}, {pp} ids x {kk} instances per GPU, "
                                           "random-init weights", "global_batch": pp * kk * world,
                               "parallelism": f"dp{world}: per-rank P x K batches, one NCCL mean all-reduce of the gradients"},
                    "tflops": r["tflops"], "loss": r["loss"], "peak_mem_gib": r["peak_mem_gib"], "clocks": clk.summary(),
                    "roofline": None, "e2e": None, "note": r["note"]}
        else:
            with ClockSampler(local) as clk:
                with clk.window():
                    r = run_retrieval(args, world, rank, local)
            line = {"metric": r["metric"], "value": r["value"], "unit": r["unit"], "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f16x3 (fp32-equivalent split)", "data": "synthetic",
                    "config": {"workload": "3368 query x 15913 gallery x 2048-d L2 top-100 + CMC/mAP (Market1501 shape)"},
                    "roofline": r["roofline"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches_per_step"] * args.steps,
                    "clocks": clk.summary(), "mAP": r["mAP"]}
            if rank == 0 and not args.no_secondary:
                rv, rdt = cpu_retrieval(128)
                line["cpu_baseline"] = {"value": rv, "unit": "pairs/s", "cores": _host_threads(), "kind": "port",
                                        "sample": f"128 of {RET_Q} queries x {RET_G} gallery, {rdt:.1f} s"}
        if rank == 0:
            print(json.dumps(line))
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
