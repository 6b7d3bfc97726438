"""Query x gallery retrieval engine: distance planes, full matrices, streamed top-k and
streamed CMC / mAP, single GPU or gallery-sharded over ranks.

Host-side mirror of the reference's retrieval path (utils/reid_metric.py:112-136,
utils/eval_reid.py:25-92, inference/get_similar.py:104-128) on top of the C ABI in
include/ctl_b200.h.  torch is used for device memory, streams and torch.distributed only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native as N

K_LIST = (1, 5, 10, 20, 50)  # utils/eval_reid.py:15


@dataclass
class Planes:
    """Opaque operand buffer of the distance kernel (see ctl_planes_build)."""

    buf: torch.Tensor
    n: int
    d: int
    flags: int
    # rows stored in another order than the caller's (build_planes(order=)): stored row i == caller's row order[i].
    # Results are always reported in the CALLER's indexing (gallery: ctl_pass_desc.g_index_map; queries: un-permuted).
    order: Optional[torch.Tensor] = None        # int32 on the device
    order_host: Optional[np.ndarray] = None

    @property
    def ptr(self):
        return self.buf.data_ptr()


def _flags(dist: str, normalize: bool) -> int:
    if dist not in ("euclidean", "cosine", "euclidean_sqrt"):
        raise KeyError(dist)
    f = N.CTL_DIST_COSINE if dist == "cosine" else N.CTL_DIST_EUCLIDEAN
    if dist == "euclidean_sqrt":
        f |= N.CTL_DIST_SQRT
    if normalize:
        f |= N.CTL_FLAG_NORMALIZE
    return f


def pid_order(pids) -> np.ndarray:
    """Stable order that sorts rows by identity.  With BOTH operands stored in this order almost every 128 x 128 tile
    of the distance GEMM pairs rows of disjoint identity ranges: such a tile holds no positive, so the pass that collects
    the positives (and the top-k threshold, which any subset of the gallery bounds) does not run it
    (ctl_pass_desc.tile_list) -- the results (indices, distances, ranks, AP) are bit-identical to the unsorted run."""
    return np.argsort(np.asarray(pids), kind="stable")


def pid_order_pays(nq: int, ng: int) -> bool:
    """topk_and_eval: identity-ordered planes trade a cheaper pass 1 (tile list: ~30 % of the matrix) for a lumpier pass 2
    (the positives and nearest rows of a query tile sit in a few gallery tiles, and the looser subset threshold lengthens
    the candidate lists).  Measured on a B200: 3368 x 15913 -> 1.41 ms sorted vs 1.30 ms in caller order;
    16384 x 25000 -> 7.9 ms vs 9.8 ms.  evaluate_streamed (no candidates) gains at both sizes."""
    return int(nq) * int(ng) >= 150_000_000


def build_planes(x: torch.Tensor, dist: str = "euclidean", normalize: bool = False, order=None) -> Planes:
    """`order` (optional, a permutation of the rows, e.g. pid_order(pids)): the planes hold x[order]."""
    N.require_cuda(x)
    if x.dim() != 2:
        raise ValueError(f"expected [n, d] features, got {tuple(x.shape)}")
    x = x.detach().float()
    order_dev = order_host = None
    if order is not None:
        order_host = np.ascontiguousarray(np.asarray(order, dtype=np.int64))
        if order_host.shape != (x.shape[0],):
            raise ValueError("order must be a permutation of the rows")
        o64 = torch.from_numpy(order_host).to(x.device, non_blocking=True)
        x = x.index_select(0, o64)
        order_dev = o64.to(torch.int32)
    x = x.contiguous()
    n, d = x.shape
    flags = _flags(dist, normalize)
    L = N.lib()
    buf = torch.empty(L.ctl_planes_bytes(n, d), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        N.check(L.ctl_planes_build(x.data_ptr(), n, d, flags, buf.data_ptr(), N.stream_ptr()))
    return Planes(buf, n, d, flags, order_dev, order_host)


class PlaneCache:
    """Operand planes of feature matrices that do not change between evaluations (a fixed gallery of `embeddings.npy`
    searched by many query batches, inference/get_similar.py:104-128; or one validation set evaluated under several
    settings): keyed on the tensor's storage pointer, shape, in-place version counter and the distance flags, so a
    modified tensor is re-packed.  Holds at most `capacity` plane buffers."""

    def __init__(self, capacity: int = 4):
        self.capacity = capacity
        self._items = {}

    def get(self, x: torch.Tensor, dist: str = "euclidean", normalize: bool = False, order=None) -> Planes:
        okey = None if order is None else hash(np.ascontiguousarray(np.asarray(order, dtype=np.int64)).tobytes())
        key = (x.data_ptr(), tuple(x.shape), x._version, str(x.device), dist, normalize, okey)
        p = self._items.get(key)
        if p is None:
            if len(self._items) >= self.capacity:
                self._items.pop(next(iter(self._items)))
            p = self._items[key] = build_planes(x, dist, normalize, order)
        return p


def dist_matrix(x: torch.Tensor, y: torch.Tensor, dist: str = "euclidean", normalize: bool = False) -> torch.Tensor:
    """get_euclidean / get_cosine (utils/reid_metric.py:25-59): the full [m, n] matrix."""
    qp, gp = build_planes(x, dist, normalize), build_planes(y, dist, normalize)
    if qp.d != gp.d:
        raise ValueError("feature dims differ")
    out = torch.empty(qp.n, gp.n, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib().ctl_dist_matrix(qp.ptr, qp.n, gp.ptr, gp.n, qp.d, qp.flags, out.data_ptr(), gp.n, N.stream_ptr()))
    return out


def topk(qp: Planes, gp: Planes, k: int, g_index_offset: int = 0, exact_threshold_pass: bool = False):
    """k nearest gallery rows per query in ascending (distance, index) order.
    Returns (idx int64 [nq, k], dist float32 [nq, k], overflow flag) on the device.  The threshold pass runs every s-th
    gallery tile only (any subset of the gallery bounds the k-th distance from above; ctl_dist_subset_stride) unless
    `exact_threshold_pass`: identical results, the subset only trades longer candidate lists for ~2/3 of that pass."""
    if qp.order is not None or gp.order is not None:
        raise ValueError("topk() takes planes in the caller's row order (use topk_and_eval for pid-sorted planes)")
    L = N.lib()
    k = int(min(k, gp.n))
    dev = qp.buf.device
    idx = torch.empty(qp.n, k, dtype=torch.int64, device=dev)
    dst = torch.empty(qp.n, k, dtype=torch.float32, device=dev)
    ovf = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = L.ctl_topk_workspace_bytes(qp.n, gp.n, k)
    if ws_bytes == 0:
        N.check(-3)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        flags = qp.flags | (N.CTL_FLAG_EXACT_PASS if exact_threshold_pass else 0)
        N.check(L.ctl_l2_topk(qp.ptr, qp.n, gp.ptr, gp.n, qp.d, flags, k, g_index_offset, idx.data_ptr(),
                              dst.data_ptr(), ovf.data_ptr(), ws.data_ptr(), ws_bytes, N.stream_ptr()))
    return idx, dst, ovf


def topk_dense(q: torch.Tensor, g: torch.Tensor, k: int, dist: str = "euclidean", normalize: bool = False,
               chunk_bytes: int = 1 << 30):
    """The k smallest distances per query from MATERIALISED rows of the distance matrix (the reference's own
    dist + argsort + `[:, :topk]`, inference/get_similar.py:104-128), chunked over queries so the matrix slice stays under
    `chunk_bytes`: the general path for every (gallery size, k) the streamed kernel's plan does not cover and for
    degenerate inputs (hundreds of exact ties at the k-th distance).  Same canonical ascending (distance, index) order:
    rows of packed integer keys are sorted."""
    nq, ng = q.shape[0], g.shape[0]
    k = int(min(k, ng))
    gp = build_planes(g, dist, normalize)
    rows = max(1, int(chunk_bytes // (12 * ng)))
    col = torch.arange(ng, device=q.device, dtype=torch.int64)[None, :]
    idx = torch.empty(nq, k, dtype=torch.int64, device=q.device)
    dst = torch.empty(nq, k, dtype=torch.float32, device=q.device)
    flip = -(1 << 63)  # unsigned key order == signed order of key ^ 2^63
    for lo in range(0, nq, rows):
        qp = build_planes(q[lo:lo + rows], dist, normalize)
        d = torch.empty(qp.n, ng, dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            N.check(N.lib().ctl_dist_matrix(qp.ptr, qp.n, gp.ptr, ng, qp.d, qp.flags, d.data_ptr(), ng, N.stream_ptr()))
        keys = (pack_keys(d, col.expand(qp.n, ng)) ^ flip).topk(k, dim=1, largest=False, sorted=True).values ^ flip
        idx[lo:lo + rows] = keys & 0xFFFFFFFF
        dst[lo:lo + rows] = torch.gather(d, 1, idx[lo:lo + rows])
    return idx, dst


def topk_similar(q: torch.Tensor, g: torch.Tensor, k: int = 100, dist: str = "euclidean", normalize: bool = False):
    """inference/get_similar.py:104-128 without the distance matrix: (indices, distances).  The reference's
    `argsort[:, :topk]` works for every topk; so does this: the streamed two-pass kernel where its plan applies
    (ctl_topk_plan: k <= ceil(ng / 16) merged column groups, candidate capacity <= 16384), otherwise -- and when more rows
    than the candidate capacity tie at the threshold -- the materialised path `topk_dense`."""
    import ctypes as C

    k = int(min(k, g.shape[0]))
    a, b, c, d_ = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    rc = N.lib().ctl_topk_plan(g.shape[0], k, C.byref(a), C.byref(b), C.byref(c), C.byref(d_))
    if rc == 0 and N.lib().ctl_topk_workspace_bytes(q.shape[0], g.shape[0], k) > 0:
        qp, gp = build_planes(q, dist, normalize), build_planes(g, dist, normalize)
        for exact in (False, True):  # a candidate overflow of the bounded threshold pass: once more with the exact one
            idx, dst, ovf = topk(qp, gp, k, exact_threshold_pass=exact)
            if int(ovf.item()) == 0:
                return idx, dst
    return topk_dense(q, g, k, dist, normalize)


# ----------------------------------------------------------------------------------------
# identities -> dense int32 pids / camera indices / camera bit masks
# ----------------------------------------------------------------------------------------


def encode_identities(q_pids, g_pids, q_camids, g_camids, respect_camids: bool):
    """Host-side re-labelling for the eval kernels (utils/eval_reid.py:52-59): pids -> dense
    int32; cameras -> dense index < 64; gallery cameras -> bit mask (one bit, or -- with
    respect_camids -- the set of cameras a centroid was built from)."""
    q_pids = np.asarray(q_pids)
    g_pids = np.asarray(g_pids)
    uniq, inv = np.unique(np.concatenate([q_pids, g_pids]), return_inverse=True)
    qp = inv[: len(q_pids)].astype(np.int32)
    gp = inv[len(q_pids):].astype(np.int32)
    n_g = len(g_pids)
    if respect_camids:
        q_cam_vals = [c[0] if isinstance(c, (list, tuple, np.ndarray)) else c for c in q_camids]
        g_sets = [list(np.atleast_1d(c)) for c in list(g_camids)[:n_g]]
        cams = sorted(set(q_cam_vals) | {c for s in g_sets for c in s})
        if len(cams) > 64:
            raise NotImplementedError(f"{len(cams)} distinct cameras; the junk filter packs camera sets into 64 bits")
        cam_index = {c: i for i, c in enumerate(cams)}
        qc = np.asarray([cam_index[c] for c in q_cam_vals], dtype=np.int32)
        gm = np.zeros(n_g, dtype=np.uint64)
        for i, s in enumerate(g_sets):
            m = 0
            for c in s:
                m |= 1 << cam_index[c]
            gm[i] = m
    else:
        q_cam = np.asarray(q_camids)
        g_cam = np.asarray(g_camids)[:n_g]  # may be over-long (bases.py:255-260 quirk)
        cams, inv_c = np.unique(np.concatenate([q_cam, g_cam]), return_inverse=True)
        if len(cams) > 64:
            raise NotImplementedError(f"{len(cams)} distinct cameras; the junk filter packs camera sets into 64 bits")
        qc = inv_c[: len(q_cam)].astype(np.int32)
        gm = np.uint64(1) << inv_c[len(q_cam):].astype(np.uint64)
    # upper bound of positives per query: the largest pid group in the gallery
    max_pos = int(np.bincount(gp).max()) if n_g else 1
    return qp, qc, gp, gm, max(1, max_pos)


@dataclass
class EncodedIds:
    """Device-resident identity arrays of one (query set, gallery set): encode once per validation set
    (identities do not change between epochs) and pass as `ids=` to skip the host re-labelling."""

    q_pid: torch.Tensor
    q_cam: torch.Tensor
    g_pid: torch.Tensor
    g_mask: torch.Tensor
    max_pos: int


def encode_ids(q_pids, g_pids, q_camids, g_camids, respect_camids: bool, device, global_labels: bool = False,
               q_order=None, g_order=None) -> EncodedIds:
    """`q_order` / `g_order`: the row orders of the planes these identities go with (Planes.order_host); the inputs are in
    the caller's order."""
    if global_labels:
        arrs = _encode_identities_global(q_pids, g_pids, q_camids, g_camids)
    else:
        arrs = encode_identities(q_pids, g_pids, q_camids, g_camids, respect_camids)
    if q_order is not None:
        arrs = (arrs[0][q_order], arrs[1][q_order]) + tuple(arrs[2:])
    if g_order is not None:
        arrs = tuple(arrs[:2]) + (arrs[2][g_order], arrs[3][g_order], arrs[4])

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)

    return EncodedIds(to_dev(arrs[0]), to_dev(arrs[1]), to_dev(arrs[2]), to_dev(arrs[3].view(np.int64)), arrs[4])


class EvalResult:
    """eval_func's outputs (utils/eval_reid.py:86-92).  `cmc`, `mAP`, `all_topk` are reduced from ONE packed read-back
    of (AP, first hit rank, positive count) per query; `single_performance` is assembled and the full `ranks` matrix
    ([nq, max_pos] 1-based kept ranks of every positive, -1 padded) is copied from the device on first access."""

    def __init__(self, cmc, mAP, all_topk, valid_idx, aps, q_pids, ranks_dev):
        self.cmc = cmc                    # float32 [max_rank]
        self.mAP = mAP
        self.all_topk = all_topk          # float64 [5]
        self._valid_idx, self._aps, self._q_pids, self._ranks_dev = valid_idx, aps, q_pids, ranks_dev
        self._ranks = None

    @property
    def single_performance(self) -> np.ndarray:  # [n_valid, 3] (q_idx, q_pid, AP)
        q = self._valid_idx
        return np.column_stack((q.astype(np.float64), np.asarray(self._q_pids)[q].astype(np.float64), self._aps))

    @property
    def ranks(self) -> np.ndarray:
        if self._ranks is None:
            r = self._ranks_dev
            self._ranks = r.cpu().numpy() if torch.is_tensor(r) else np.asarray(r)
        return self._ranks


def _aggregate(ranks, ap: np.ndarray, n_pos: np.ndarray, q_pids, num_g: int, max_rank: int, first=None) -> EvalResult:
    """The reductions at the end of eval_func (utils/eval_reid.py:86-92) from per-query results: O(nq) on the host
    (a histogram of the first-hit ranks gives the whole CMC curve), bit-identical to the reference's float32 / float64
    reductions.  `ranks` may stay on the device (only `first` = ranks[:, 0] is needed here)."""
    max_rank = min(max_rank, num_g)
    valid = n_pos > 0
    if not valid.any():
        raise RuntimeError("no valid query: no query identity appears in the gallery")
    if first is None:
        first = np.asarray(ranks)[:, 0]
    first = first[valid].astype(np.int64)  # ranks are sorted ascending -> the first hit
    num_valid = int(valid.sum())
    hist = np.bincount(np.minimum(first, max_rank + 1), minlength=max_rank + 2)[1: max_rank + 1]
    hits = np.cumsum(hist)                                     # queries whose first hit is at rank <= r
    cmc = hits.astype(np.float32) / np.float32(num_valid)      # float32 sum of 0/1 rows / count, as the reference
    topk = np.asarray([float(hits[k - 1]) if k <= max_rank else float(num_valid) for k in K_LIST]) / float(num_valid)
    aps = ap[valid]
    return EvalResult(cmc, float(np.mean(aps)), topk, np.nonzero(valid)[0], aps, q_pids, ranks)


def _finalize(buckets, count, nq, max_pos, ovf):
    """ctl_eval_finalize_packed (enqueue only): ranks [nq, max_pos] and the packed per-query results [nq + 1, 3] float64
    (AP, first-hit rank, #positives; last row: overflow flag) on the device."""
    dev = buckets.device
    ranks = torch.empty(nq, max_pos, dtype=torch.int32, device=dev)
    ap = torch.empty(nq, dtype=torch.float64, device=dev)
    pack = torch.empty(nq + 1, 3, dtype=torch.float64, device=dev)
    N.check(N.lib().ctl_eval_finalize_packed(buckets.data_ptr(), count.data_ptr(), nq, max_pos, ranks.data_ptr(), ap.data_ptr(),
                                             pack.data_ptr(), ovf.data_ptr(), N.stream_ptr()))
    return ranks, pack


def _unpack(h: np.ndarray, nq: int):
    """host view of `pack`: (ap, first-hit rank, #positives, overflow flag) -- float64 is exact for these integers."""
    return h[:nq, 0], h[:nq, 1].astype(np.int64), h[:nq, 2].astype(np.int32), int(h[nq, 0])


def _finalize_and_read_back(buckets, count, nq, max_pos, ovf):
    """_finalize + ONE device->host copy.  Returns (ranks on the device, ap, first, count, overflow) -- the last four on
    the host."""
    ranks, pack = _finalize(buckets, count, nq, max_pos, ovf)
    return (ranks,) + _unpack(pack.cpu().numpy(), nq)


def _tile_lists_enabled(qp: Planes, gp: Planes) -> bool:
    """Tile lists (ctl_pass_desc.tile_list) pay off when BOTH operands are stored in identity order (then few tiles can
    hold a positive); off with CTL_RETRIEVAL_TILE_LISTS=0 (bisect aid)."""
    import os

    return os.environ.get("CTL_RETRIEVAL_TILE_LISTS", "1") != "0" and qp.order is not None and gp.order is not None


def _tile_list(qp: Planes, gp: Planes, ids: "EncodedIds", keep_stride: int) -> Optional[torch.Tensor]:
    """ctl_dist_worklist: the tiles that can hold a positive (+ every keep_stride-th gallery tile for the threshold).
    None when the problem is beyond the list builder (the pass then runs every tile)."""
    L = N.lib()
    if ((qp.n + 127) // 128) * ((gp.n + 127) // 128) > (1 << 20) or (qp.n + 127) // 128 + (gp.n + 127) // 128 > 5632:
        return None
    work = torch.empty(L.ctl_dist_worklist_bytes(qp.n, gp.n) // 4, dtype=torch.int32, device=qp.buf.device)
    N.check(L.ctl_dist_worklist(ids.q_pid.data_ptr(), qp.n, ids.g_pid.data_ptr(), gp.n, int(keep_stride), work.data_ptr(),
                                N.stream_ptr()))
    return work


def _g_index_map(gp: Planes, g_index_offset: int) -> Optional[torch.Tensor]:
    """int32 map stored gallery row -> index reported in the results (None: row + g_index_offset)."""
    if gp.order is None:
        return None
    if g_index_offset + gp.n >= (1 << 31):
        raise NotImplementedError("re-ordered gallery planes report int32 indices")
    return gp.order if g_index_offset == 0 else (gp.order + int(g_index_offset)).to(torch.int32)


def _query_inverse(qp: Planes):
    """(device int64, host) inverse of the query row order: results[inv] are in the caller's order."""
    if qp.order is None:
        return None, None
    inv = getattr(qp, "_inv", None)
    if inv is None:
        inv_h = np.empty(qp.n, dtype=np.int64)
        inv_h[qp.order_host] = np.arange(qp.n)
        inv = qp._inv = (torch.from_numpy(inv_h).to(qp.buf.device), inv_h)
    return inv


def evaluate_streamed(
    qp: Planes,
    gp: Planes,
    q_pids,
    g_pids,
    q_camids,
    g_camids,
    max_rank: int = 50,
    respect_camids: bool = False,
    g_index_offset: int = 0,
    group=None,
    total_gallery: Optional[int] = None,
    ids: "Optional[EncodedIds]" = None,
) -> EvalResult:
    """eval_func semantics (utils/eval_reid.py:25-92) straight from the features: two tensor-
    core passes (collect the positives' distances; count kept rows before each positive),
    no distance matrix, no argsort.  With `group` (torch.distributed), `gp` is this rank's
    gallery shard and g_* its identities; keys are all-gathered, buckets all-reduced.
    Identities are given in the caller's row order even when the planes were built with `order=`; a precomputed `ids`
    must have been encoded with the planes' orders (encode_ids(q_order=, g_order=))."""
    import ctypes as C

    import torch.distributed as dist

    L = N.lib()
    dev = qp.buf.device
    nq, ng = qp.n, gp.n
    world = dist.get_world_size(group) if group is not None else 1
    if ids is None:
        if world > 1 and not np.issubdtype(np.asarray(q_pids).dtype, np.integer):
            raise ValueError("sharded evaluation needs integer pids")
        # sharded: dense re-labelling must agree across ranks -> identity map instead of np.unique
        ids = encode_ids(q_pids, g_pids, q_camids, g_camids, respect_camids, dev, global_labels=world > 1,
                         q_order=qp.order_host, g_order=gp.order_host)
    d_qpid, d_qcam, d_gpid, d_gmask, max_pos_local = ids.q_pid, ids.q_cam, ids.g_pid, ids.g_mask, ids.max_pos
    max_pos = max_pos_local
    if world > 1:
        mp = torch.tensor([max_pos_local], device=dev, dtype=torch.int64)
        dist.all_reduce(mp, op=dist.ReduceOp.SUM, group=group)  # positives of a pid may spread over shards
        max_pos = int(mp.item())
    pos_keys = torch.zeros(nq, max_pos, dtype=torch.int64, device=dev)
    pos_count = torch.zeros(nq, dtype=torch.int32, device=dev)
    ovf = torch.zeros(1, dtype=torch.int32, device=dev)
    s = N.stream_ptr
    gmap = _g_index_map(gp, g_index_offset)
    idp = dict(q_pid=d_qpid.data_ptr(), q_cam=d_qcam.data_ptr(), g_pid=d_gpid.data_ptr(), g_cammask=d_gmask.data_ptr(),
               max_pos=max_pos, overflow=ovf.data_ptr(), g_index_offset=g_index_offset, g_index_map=N.ptr(gmap))
    with torch.cuda.device(dev):
        # pass 1 (collect) only wants the positives: tiles whose identity ranges are disjoint are not run -- with both
        # operands stored in pid order (build_planes(order=pid_order(..))) that is ~95 % of a Market-sized problem
        work = _tile_list(qp, gp, ids, 0) if _tile_lists_enabled(qp, gp) else None
        p1 = N.PassDesc(pos_keys=pos_keys.data_ptr(), pos_count=pos_count.data_ptr(), tile_list=N.ptr(work), **idp)
        N.check(L.ctl_dist_pass(qp.ptr, nq, gp.ptr, ng, qp.d, qp.flags, C.byref(p1), s()))
        if world > 1:
            pos_keys, pos_count = _allgather_keys(pos_keys, pos_count, max_pos, group)
        N.check(L.ctl_sort_key_rows(pos_keys.data_ptr(), pos_count.data_ptr(), nq, max_pos, s()))
        buckets = torch.zeros(nq, max_pos + 1, dtype=torch.int32, device=dev)
        p2 = N.PassDesc(thr_keys=pos_keys.data_ptr(), thr_count=pos_count.data_ptr(), buckets=buckets.data_ptr(), **idp)
        N.check(L.ctl_dist_pass(qp.ptr, nq, gp.ptr, ng, qp.d, qp.flags, C.byref(p2), s()))
        if world > 1:
            dist.all_reduce(buckets, op=dist.ReduceOp.SUM, group=group)
        ranks, ap_h, first_h, cnt_h, ovf_h = _finalize_and_read_back(buckets, pos_count, nq, max_pos, ovf)
    if ovf_h:
        raise OverflowError("positives list overflowed (max_pos too small)")
    inv_d, inv_h = _query_inverse(qp)
    if inv_h is not None:  # back to the caller's query order
        ranks, ap_h, first_h, cnt_h = ranks.index_select(0, inv_d), ap_h[inv_h], first_h[inv_h], cnt_h[inv_h]
    num_g = total_gallery if total_gallery is not None else ng
    return _aggregate(ranks, ap_h, cnt_h, np.asarray(q_pids), num_g, max_rank, first=first_h)


def _encode_identities_global(q_pids, g_pids, q_camids, g_camids):
    q_pid = np.asarray(q_pids).astype(np.int32)
    g_pid = np.asarray(g_pids).astype(np.int32)
    q_cam = np.asarray(q_camids).astype(np.int32)
    g_cam = np.asarray(g_camids).astype(np.int64)[: len(g_pid)]
    if q_cam.max(initial=0) >= 64 or g_cam.max(initial=0) >= 64 or min(q_cam.min(initial=0), g_cam.min(initial=0)) < 0:
        raise NotImplementedError("sharded evaluation expects camera ids in [0, 64)")
    g_mask = (np.uint64(1) << g_cam.astype(np.uint64)).astype(np.uint64)
    max_pos = int(np.bincount(g_pid - g_pid.min()).max()) if len(g_pid) else 1
    return q_pid, q_cam, g_pid, g_mask, max(1, max_pos)


def _allgather_keys(pos_keys, pos_count, max_pos, group):
    """Concatenates every rank's positives per query (ragged, packed to the left)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    keys_all = [torch.empty_like(pos_keys) for _ in range(world)]
    cnt_all = [torch.empty_like(pos_count) for _ in range(world)]
    dist.all_gather(keys_all, pos_keys, group=group)
    dist.all_gather(cnt_all, pos_count, group=group)
    nq = pos_keys.shape[0]
    out = torch.zeros_like(pos_keys)
    total = torch.zeros_like(pos_count)
    col = torch.arange(max_pos, device=pos_keys.device)[None, :]
    for kr, cr in zip(keys_all, cnt_all):
        valid = col < cr[:, None]
        dest = (total[:, None] + col).clamp(max=max_pos - 1)
        rows = torch.arange(nq, device=pos_keys.device)[:, None].expand_as(dest)
        out[rows[valid], dest[valid].long()] = kr[valid]
        total = total + cr
    return out, total


def merge_topk(idx_list: Sequence[torch.Tensor], dist_list: Sequence[torch.Tensor], k: int):
    """k-way merge of per-shard (ascending) top-k lists under the canonical (distance, index)
    order -- deterministic regardless of world size.  CUDA inputs are merged by the native packed-key row sort
    (one launch, integer-exact); host tensors (the gloo tests) by two stable argsorts."""
    idx = torch.cat(list(idx_list), 1)
    dst = torch.cat(list(dist_list), 1)
    if idx.is_cuda:
        return merge_topk_keys(pack_keys(dst, idx), k)
    # sort by index first (stable), then by distance (stable): lexicographic (distance, index)
    o1 = torch.argsort(idx, dim=1, stable=True)
    idx, dst = idx.gather(1, o1), dst.gather(1, o1)
    o2 = torch.argsort(dst, dim=1, stable=True)
    return idx.gather(1, o2)[:, :k], dst.gather(1, o2)[:, :k]


def pack_keys(dst: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """(distance fp32, index) -> the kernels' uint64 key (carried as int64): orderable(fp32) << 32 | index, whose
    UNSIGNED integer order is the canonical ascending (distance, index) order (ctl_key_encode)."""
    b = dst.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    o = torch.where(b >= 0x80000000, b ^ 0xFFFFFFFF, b | 0x80000000)  # sign-magnitude float bits -> unsigned order
    return (o << 32) | (idx.to(torch.int64) & 0xFFFFFFFF)


def merge_topk_keys(keys: torch.Tensor, k: int):
    """keys: int64 [nq, m] packed (distance, index) keys in any order (m >= k) -> the k smallest per row as
    (idx int64 [nq, k], dist float32 [nq, k]): ctl_sort_key_rows + ctl_topk_emit."""
    L = N.lib()
    keys = keys.contiguous()
    nq, m = keys.shape
    dev = keys.device
    counts = torch.full((nq,), m, dtype=torch.int32, device=dev)
    idx = torch.empty(nq, k, dtype=torch.int64, device=dev)
    dst = torch.empty(nq, k, dtype=torch.float32, device=dev)
    ovf = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(L.ctl_sort_key_rows(keys.data_ptr(), counts.data_ptr(), nq, m, N.stream_ptr()))
        N.check(L.ctl_topk_emit(keys.data_ptr(), counts.data_ptr(), nq, m, k, idx.data_ptr(), dst.data_ptr(),
                                ovf.data_ptr(), N.stream_ptr()))
    return idx, dst


def topk_sharded(q_local: torch.Tensor, g_local: torch.Tensor, k: int, g_index_offset: int, group,
                 dist: str = "euclidean", normalize: bool = False):
    """BASELINE config 5: queries sharded by rank are all-gathered once (NCCL), the gallery
    stays sharded; every rank returns the merged global top-k of ALL queries."""
    import torch.distributed as tdist

    world = tdist.get_world_size(group)
    q_all = [torch.empty_like(q_local) for _ in range(world)]
    tdist.all_gather(q_all, q_local.contiguous(), group=group)
    q = torch.cat(q_all, 0)
    qp, gp = build_planes(q, dist, normalize), build_planes(g_local, dist, normalize)
    idx, dst, ovf = topk(qp, gp, k, g_index_offset)
    idx_all = [torch.empty_like(idx) for _ in range(world)]
    dst_all = [torch.empty_like(dst) for _ in range(world)]
    tdist.all_gather(idx_all, idx, group=group)
    tdist.all_gather(dst_all, dst, group=group)
    tdist.all_reduce(ovf, group=group)
    if int(ovf.item()) != 0:
        raise OverflowError("top-k candidate capacity exceeded on some rank")
    return merge_topk(idx_all, dst_all, k)


def topk_and_eval(qp: Planes, gp: Planes, k: int, q_pids, g_pids, q_camids, g_camids, max_rank: int = 50,
                  respect_camids: bool = False, ids: "Optional[EncodedIds]" = None, tile_lists: Optional[bool] = None):
    """BASELINE config 3 in TWO tensor-core passes: per-query top-k (ascending (distance, index))
    AND eval_func's CMC / mAP, neither materialising the distance matrix.
      pass 1: 16-column group minima (-> tau) + the positives' distances
      pass 2: candidates <= tau + kept rows before each positive
    Pass 1 does not need the whole matrix: the positives sit in the tiles whose query / gallery identity ranges
    intersect, and the k-th smallest group minimum of ANY subset of the gallery bounds the k-th distance from above.
    With both operands stored in pid order (build_planes(order=pid_order(pids))) pass 1 therefore runs a tile list
    (ctl_dist_worklist: the few tiles that can hold a positive + every s-th gallery tile, ~30 % of the matrix); the
    looser tau only lengthens the candidate lists of the exact pass 2, so indices, distances, ranks and AP are
    bit-identical to the full run (`tile_lists=False`).
    Identities are given in the caller's row order; a precomputed `ids` must carry the planes' orders
    (encode_ids(q_order=qp.order_host, g_order=gp.order_host)).
    Returns (idx [nq,k] int64, dist [nq,k] float32 on the device, EvalResult), all in the caller's indexing."""
    dev = qp.buf.device
    if ids is None:
        ids = encode_ids(q_pids, g_pids, q_camids, g_camids, respect_camids, dev, q_order=qp.order_host,
                         g_order=gp.order_host)
    if tile_lists is None:
        tile_lists = _tile_lists_enabled(qp, gp)
    with torch.cuda.device(dev):
        out = _topk_and_eval_enqueue(qp, gp, k, ids, tile_lists)
        h = out["pack"].cpu().numpy()
    return _topk_and_eval_finish(out, h, qp, gp, k, q_pids, g_pids, q_camids, g_camids, max_rank, respect_camids, ids)


def _topk_and_eval_enqueue(qp: Planes, gp: Planes, k: int, ids: "EncodedIds", tile_lists: bool):
    """The launch sequence of topk_and_eval, no host synchronisation (capturable in a CUDA graph): returns the device
    tensors {idx, dst, ranks, pack} in the planes' row order and whether pass 1 ran a threshold subset."""
    import ctypes as C

    L = N.lib()
    dev = qp.buf.device
    nq, ng = qp.n, gp.n
    k = int(min(k, ng))
    emit_all, n_groups, merge, cap = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    N.check(L.ctl_topk_plan(ng, k, C.byref(emit_all), C.byref(n_groups), C.byref(merge), C.byref(cap)))
    d_qpid, d_qcam, d_gpid, d_gmask, max_pos = ids.q_pid, ids.q_cam, ids.g_pid, ids.g_mask, ids.max_pos
    gmin = torch.empty(nq, n_groups.value, dtype=torch.float32, device=dev)
    tau = torch.empty(nq, dtype=torch.float32, device=dev)
    cand = torch.empty(nq, cap.value, dtype=torch.int64, device=dev)
    zeros = torch.zeros(2 * nq + 1, dtype=torch.int32, device=dev)
    cand_count, pos_count, ovf = zeros[:nq], zeros[nq: 2 * nq], zeros[2 * nq:]
    pos_keys = torch.empty(nq, max_pos, dtype=torch.int64, device=dev)
    buckets = torch.zeros(nq, max_pos + 1, dtype=torch.int32, device=dev)
    idx = torch.empty(nq, k, dtype=torch.int64, device=dev)
    dst = torch.empty(nq, k, dtype=torch.float32, device=dev)
    s = N.stream_ptr
    gmap = _g_index_map(gp, 0)
    idp = dict(q_pid=d_qpid.data_ptr(), q_cam=d_qcam.data_ptr(), g_pid=d_gpid.data_ptr(),
               g_cammask=d_gmask.data_ptr(), max_pos=max_pos, overflow=ovf.data_ptr(), g_index_map=N.ptr(gmap))
    p1 = N.PassDesc(pos_keys=pos_keys.data_ptr(), pos_count=pos_count.data_ptr(), **idp)
    if not emit_all.value:
        p1.gmin = gmin.data_ptr()
    work = None
    if tile_lists:
        # (small gallery, tau = +inf: pass 1 only collects -> stride 0, just the tiles that can hold a positive)
        stride = 0 if emit_all.value else L.ctl_dist_subset_stride(ng, k)
        if emit_all.value or stride > 1:
            work = _tile_list(qp, gp, ids, stride)
    if work is not None:
        p1.tile_list = work.data_ptr()
        if not emit_all.value:
            N.check(L.ctl_fill_f32(gmin.data_ptr(), gmin.numel(), float("inf"), s()))  # groups of tiles not run
    N.check(L.ctl_dist_pass(qp.ptr, nq, gp.ptr, ng, qp.d, qp.flags, C.byref(p1), s()))
    if emit_all.value:
        N.check(L.ctl_fill_f32(tau.data_ptr(), nq, float("inf"), s()))
    else:
        N.check(L.ctl_select_tau(gmin.data_ptr(), nq, n_groups.value, merge.value, k, tau.data_ptr(), s()))
    N.check(L.ctl_sort_key_rows(pos_keys.data_ptr(), pos_count.data_ptr(), nq, max_pos, s()))
    p2 = N.PassDesc(tau=tau.data_ptr(), cand_keys=cand.data_ptr(), cand_count=cand_count.data_ptr(),
                    cand_cap=cap.value, thr_keys=pos_keys.data_ptr(), thr_count=pos_count.data_ptr(),
                    buckets=buckets.data_ptr(), **idp)
    N.check(L.ctl_dist_pass(qp.ptr, nq, gp.ptr, ng, qp.d, qp.flags, C.byref(p2), s()))
    N.check(L.ctl_sort_key_rows(cand.data_ptr(), cand_count.data_ptr(), nq, cap.value, s()))
    N.check(L.ctl_topk_emit(cand.data_ptr(), cand_count.data_ptr(), nq, cap.value, k, idx.data_ptr(),
                            dst.data_ptr(), ovf.data_ptr(), s()))
    ranks, pack = _finalize(buckets, pos_count, nq, max_pos, ovf)
    return {"idx": idx, "dst": dst, "ranks": ranks, "pack": pack, "subset": work is not None and not emit_all.value,
            "keep": (gmin, tau, cand, zeros, pos_keys, buckets, work, gmap)}


def _topk_and_eval_finish(out, h, qp, gp, k, q_pids, g_pids, q_camids, g_camids, max_rank, respect_camids, ids):
    """host half of topk_and_eval: overflow handling, back to the caller's query order, eval_func's final reductions."""
    nq, ng = qp.n, gp.n
    idx, dst, ranks = out["idx"], out["dst"], out["ranks"]
    ap_h, first_h, cnt_h, ovf_h = _unpack(h, nq)
    if ovf_h and out["subset"]:
        # the looser threshold let more rows through than the candidate list holds: threshold from every tile
        return topk_and_eval(qp, gp, k, q_pids, g_pids, q_camids, g_camids, max_rank, respect_camids, ids, tile_lists=False)
    if ovf_h:
        raise OverflowError("a device-side list overflowed (exact ties at the k-th distance, or max_pos)")
    inv_d, inv_h = _query_inverse(qp)
    if inv_h is not None:  # back to the caller's query order
        idx, dst, ranks = idx.index_select(0, inv_d), dst.index_select(0, inv_d), ranks.index_select(0, inv_d)
        ap_h, first_h, cnt_h = ap_h[inv_h], first_h[inv_h], cnt_h[inv_h]
    return idx, dst, _aggregate(ranks, ap_h, cnt_h, np.asarray(q_pids), ng, max_rank, first=first_h)


class TopkEvalSession:
    """topk_and_eval for ONE validation set evaluated again and again (a resident gallery searched by successive query
    batches, inference/get_similar.py:104-128; the per-epoch validation of train_ctl_model.py): the ~12 launches of the step
    (query planes, both tensor-core passes, selection, sorts, emit, finalize, the packed device->host copy) are captured
    ONCE in a CUDA graph over static buffers and replayed -- no per-step allocation, descriptor encoding or launch gaps.
    Results are those of topk_and_eval, bit for bit (tests/test_retrieval_gpu.py)."""

    def __init__(self, gallery: torch.Tensor, num_query: int, k: int, q_pids, g_pids, q_camids, g_camids, max_rank: int = 50,
                 respect_camids: bool = False, dist: str = "euclidean", normalize: bool = False):
        N.require_cuda(gallery)
        dev = gallery.device
        self.dev, self.k, self.max_rank, self.respect = dev, int(k), max_rank, respect_camids
        self.args = (q_pids, g_pids, q_camids, g_camids)
        self.dist, self.normalize = dist, normalize
        self.gp = build_planes(gallery, dist, normalize)
        self.ids = encode_ids(q_pids, g_pids, q_camids, g_camids, respect_camids, dev)
        self.q = torch.empty(num_query, gallery.shape[1], dtype=torch.float32, device=dev)  # static input of the graph
        self.host = torch.empty(num_query + 1, 3, dtype=torch.float64).pin_memory()
        self.done = torch.cuda.Event()

        def enqueue():
            qp = build_planes(self.q, dist, normalize)
            out = _topk_and_eval_enqueue(qp, self.gp, self.k, self.ids, False)
            self.host.copy_(out["pack"], non_blocking=True)
            out["qp"] = qp
            return out

        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # eager warm-up: function attributes, allocator pools
                self.q.zero_()
                for _ in range(2):
                    enqueue()
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = enqueue()

    def __call__(self, q: torch.Tensor):
        """q: [num_query, d] features on the device -> (idx, dist, EvalResult) like topk_and_eval.  The returned device
        tensors are the graph's static outputs: valid until the next call."""
        N.require_cuda(q)
        with torch.cuda.device(self.dev):
            self.q.copy_(q, non_blocking=True)
            self.graph.replay()
            self.done.record()
            self.done.synchronize()
        out = self.out
        return _topk_and_eval_finish(out, self.host.numpy(), out["qp"], self.gp, self.k, *self.args, self.max_rank,
                                     self.respect, self.ids)


def encode_ids_sharded(q_pids, g_pids_local, q_camids, g_camids_local, device, group, q_order=None,
                       g_order=None) -> EncodedIds:
    """Identity arrays of (all queries, THIS rank's gallery shard) for topk_and_eval_sharded: raw integer pids (every
    rank must agree on the labelling, so no np.unique), camera ids in [0, 64); `max_pos` = the largest number of
    same-pid rows of any shard (one MAX all-reduce, done once per validation set)."""
    import torch.distributed as dist

    ids = encode_ids(q_pids, g_pids_local, q_camids, g_camids_local, False, device, global_labels=True, q_order=q_order,
                     g_order=g_order)
    mp = torch.tensor([ids.max_pos], device=device, dtype=torch.int64)
    dist.all_reduce(mp, op=dist.ReduceOp.MAX, group=group)
    ids.max_pos = int(mp.item())
    return ids


def topk_and_eval_sharded(qp: Planes, gp_local: Planes, k: int, ids: EncodedIds, q_pids, g_index_offset: int,
                          total_gallery: int, group, max_rank: int = 50, tile_lists: Optional[bool] = None):
    """BASELINE config 5: topk_and_eval with the GALLERY AXIS SHARDED over the ranks of `group` (queries replicated:
    all-gather them once before building `qp`).  Every rank runs the two tensor-core passes over its own shard; the
    exchange steps are (utils/reid_metric.py:112-136 + utils/eval_reid.py:25-92 semantics, bit-identical to one GPU):
      after pass 1: all-gather of the positives' (distance, index) keys [nq, max_pos] -> one sorted threshold list
      after pass 2: all-reduce(sum) of the integer bucket counts; all-gather of each rank's k best packed keys and a
                    k-way merge by integer key order (world-size independent).
    Returns (idx [nq, k] global gallery rows, dist [nq, k], EvalResult) on every rank."""
    import ctypes as C

    import torch.distributed as dist

    L = N.lib()
    dev = qp.buf.device
    world = dist.get_world_size(group)
    nq, ng = qp.n, gp_local.n
    k_loc = int(min(k, ng))
    emit_all, n_groups, merge, cap = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    N.check(L.ctl_topk_plan(ng, k_loc, C.byref(emit_all), C.byref(n_groups), C.byref(merge), C.byref(cap)))
    mp_l = ids.max_pos            # per-shard capacity (same on every rank)
    mp = mp_l * world             # capacity of the merged threshold list
    gmin = torch.empty(nq, n_groups.value, dtype=torch.float32, device=dev)
    tau = torch.empty(nq, dtype=torch.float32, device=dev)
    cand = torch.empty(nq, cap.value, dtype=torch.int64, device=dev)
    zeros = torch.zeros(2 * nq + 1, dtype=torch.int32, device=dev)
    cand_count, pos_count, ovf = zeros[:nq], zeros[nq: 2 * nq], zeros[2 * nq:]
    pos_keys = torch.empty(nq, mp_l, dtype=torch.int64, device=dev)
    buckets = torch.zeros(nq, mp + 1, dtype=torch.int32, device=dev)
    s = N.stream_ptr
    if tile_lists is None:
        tile_lists = _tile_lists_enabled(qp, gp_local)
    gmap = _g_index_map(gp_local, g_index_offset)  # pid-sorted shard: keys carry the GLOBAL gallery row
    idp = dict(q_pid=ids.q_pid.data_ptr(), q_cam=ids.q_cam.data_ptr(), g_pid=ids.g_pid.data_ptr(),
               g_cammask=ids.g_mask.data_ptr(), overflow=ovf.data_ptr(), g_index_offset=g_index_offset,
               g_index_map=N.ptr(gmap))
    with torch.cuda.device(dev):
        p1 = N.PassDesc(pos_keys=pos_keys.data_ptr(), pos_count=pos_count.data_ptr(), max_pos=mp_l, **idp)
        if not emit_all.value:
            p1.gmin = gmin.data_ptr()
        work = None
        if tile_lists:  # pid-sorted shard: pass 1 runs the tiles that can hold a positive + a subset for tau (topk_and_eval)
            stride = 0 if emit_all.value else L.ctl_dist_subset_stride(ng, k_loc)
            if emit_all.value or stride > 1:
                work = _tile_list(qp, gp_local, ids, stride)
        if work is not None:
            p1.tile_list = work.data_ptr()
            if not emit_all.value:
                N.check(L.ctl_fill_f32(gmin.data_ptr(), gmin.numel(), float("inf"), s()))
        N.check(L.ctl_dist_pass(qp.ptr, nq, gp_local.ptr, ng, qp.d, qp.flags, C.byref(p1), s()))
        if emit_all.value:
            N.check(L.ctl_fill_f32(tau.data_ptr(), nq, float("inf"), s()))
        else:
            N.check(L.ctl_select_tau(gmin.data_ptr(), nq, n_groups.value, merge.value, k_loc, tau.data_ptr(), s()))
        # exchange 1: every rank's positives, unused slots = the largest key, so ONE row sort packs and orders them
        col = torch.arange(mp_l, device=dev)[None, :]
        masked = torch.where(col < pos_count[:, None].clamp(max=mp_l), pos_keys, torch.full_like(pos_keys, -1))
        g_keys = torch.empty(world, nq, mp_l, dtype=torch.int64, device=dev)
        g_cnt = torch.empty(world, nq, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(g_keys, masked, group=group)
        dist.all_gather_into_tensor(g_cnt, pos_count.contiguous(), group=group)
        thr = g_keys.permute(1, 0, 2).reshape(nq, mp).contiguous()
        thr_count = g_cnt.sum(0, dtype=torch.int32)
        full = torch.full((nq,), mp, dtype=torch.int32, device=dev)
        N.check(L.ctl_sort_key_rows(thr.data_ptr(), full.data_ptr(), nq, mp, s()))
        p2 = N.PassDesc(tau=tau.data_ptr(), cand_keys=cand.data_ptr(), cand_count=cand_count.data_ptr(),
                        cand_cap=cap.value, thr_keys=thr.data_ptr(), thr_count=thr_count.data_ptr(),
                        buckets=buckets.data_ptr(), max_pos=mp, **idp)
        N.check(L.ctl_dist_pass(qp.ptr, nq, gp_local.ptr, ng, qp.d, qp.flags, C.byref(p2), s()))
        N.check(L.ctl_sort_key_rows(cand.data_ptr(), cand_count.data_ptr(), nq, cap.value, s()))
        # exchange 2: bucket counts (integers) and the k best keys of every shard
        dist.all_reduce(buckets, op=dist.ReduceOp.SUM, group=group)
        best = cand[:, :k_loc].contiguous()
        g_best = torch.empty(world, nq, k_loc, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(g_best, best, group=group)
        dist.all_reduce(ovf, op=dist.ReduceOp.MAX, group=group)
        idx, dst = merge_topk_keys(g_best.permute(1, 0, 2).reshape(nq, world * k_loc), int(min(k, world * k_loc)))
        ranks, ap_h, first_h, cnt_h, ovf_h = _finalize_and_read_back(buckets, thr_count, nq, mp, ovf)
        if ovf_h and tile_lists and not emit_all.value:  # (the flag is MAX-reduced: every rank takes this branch together)
            return topk_and_eval_sharded(qp, gp_local, k, ids, q_pids, g_index_offset, total_gallery, group, max_rank,
                                         tile_lists=False)
        inv_d, inv_h = _query_inverse(qp)
        if inv_h is not None:  # back to the caller's query order
            idx, dst, ranks = idx.index_select(0, inv_d), dst.index_select(0, inv_d), ranks.index_select(0, inv_d)
            ap_h, first_h, cnt_h = ap_h[inv_h], first_h[inv_h], cnt_h[inv_h]
    if ovf_h:
        raise OverflowError("a device-side list overflowed on some rank (exact ties at the k-th distance, or max_pos)")
    return idx, dst, _aggregate(ranks, ap_h, cnt_h, np.asarray(q_pids), total_gallery, max_rank, first=first_h)
