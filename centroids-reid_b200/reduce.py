"""Per-identity centroid builders on the segmented-mean kernel (ctl_segment_mean).

Host-side mirrors of modelling/bases.py:92-95,179-262 and
inference/inference_utils.py:147-159.  `validation_create_centroids` builds its groups ON THE DEVICE
(sort by label, segment boundaries, camera-set bit masks, one expand + filter) and reduces them
with ONE ctl_segment_mean launch over the resulting CSR description; the reference's behaviour,
quirks included, is kept (see its docstring).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N


def segment_mean(x: torch.Tensor, groups) -> torch.Tensor:
    """out[s] = mean of x[groups[s]] (row order = summation order).  x: CUDA [n, d] fp32."""
    N.require_cuda(x)
    x = x.detach().float().contiguous()
    n, d = x.shape
    indptr = np.zeros(len(groups) + 1, dtype=np.int64)
    indptr[1:] = np.cumsum([len(g) for g in groups])
    indices = np.concatenate([np.asarray(g, dtype=np.int64) for g in groups]) if len(groups) else np.zeros(0, np.int64)
    d_ptr = torch.from_numpy(indptr).to(x.device, non_blocking=True)
    d_idx = torch.from_numpy(indices).to(x.device, non_blocking=True)
    out = torch.empty(len(groups), d, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib().ctl_segment_mean(x.data_ptr(), n, d, d_ptr.data_ptr(), d_idx.data_ptr(), len(groups),
                                         out.data_ptr(), N.stream_ptr()))
    return out


def _calculate_centroids(vecs, dim=1):
    """ModelBase._calculate_centroids (modelling/bases.py:92-95): sum over `dim` / length."""
    if vecs.dim() == 2 and dim == 0:
        return segment_mean(vecs, [np.arange(vecs.shape[0])])[0]
    if vecs.dim() == 3 and dim == 1:
        b, k, d = vecs.shape
        return segment_mean(vecs.reshape(b * k, d), [np.arange(i * k, (i + 1) * k) for i in range(b)])
    raise NotImplementedError("_calculate_centroids supports [n,d] over dim 0 and [b,k,d] over dim 1")


def calculate_centroids(embeddings, pid_path_index):
    """inference/inference_utils.py:147-159 -> (centroids_arr [n_pid, d] ndarray, pids as np.str_)."""
    emb = torch.as_tensor(np.asarray(embeddings))
    emb = emb if emb.is_cuda else emb.cuda(non_blocking=True)
    pids = list(pid_path_index.keys())
    cents = segment_mean(emb, [pid_path_index[p] for p in pids])
    return cents.cpu().numpy(), np.array(pids, dtype=np.str_)


def segment_mean_csr(x: torch.Tensor, indptr: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """ctl_segment_mean on a CSR description that already lives on the device (int64 indptr [n_seg + 1], indices)."""
    N.require_cuda(x, indptr, indices)
    x = x.detach().float().contiguous()
    n, d = x.shape
    n_seg = indptr.numel() - 1
    out = torch.empty(n_seg, d, dtype=torch.float32, device=x.device)
    if n_seg == 0:
        return out
    with torch.cuda.device(x.device):
        N.check(N.lib().ctl_segment_mean(x.data_ptr(), n, d, indptr.contiguous().data_ptr(), indices.contiguous().data_ptr(),
                                         n_seg, out.data_ptr(), N.stream_ptr()))
    return out


def validation_create_centroids(embeddings, labels, camids, num_query, respect_camids=False):
    """ModelBase.validation_create_centroids (modelling/bases.py:179-262) with the GROUPING on the device.

    The reference walks the identities in a Python loop (dict of index lists, np.unique / np.where / list comprehensions
    per identity and per query camera).  Here the gallery rows are sorted once by (label, row) on the device, identities
    are the segments of that order, the "cameras seen for this identity" sets are 64-bit masks built by one scatter, the
    candidate (identity, query camera) pairs come from one `unique`, and the members of every centroid are produced by
    one expand + filter -- a CSR description handed straight to ctl_segment_mean.  Semantics kept, quirk included:
      * one centroid per identity (sorted by label), or with respect_camids one per DISTINCT set of "other cameras"
        {c in cams(identity) : c != q} over the identity's query cameras q in ascending order, first occurrence wins,
        empty sets skipped (bases.py:208-236);
      * bases.py:214 indexes the FULL camid array with gallery-relative indices: the camera of gallery row i is taken
        from camids[i], not camids[num_query + i];
      * without respect_camids the dummy camids are 0 for queries, 1 for centroids, and the gallery part is sized from the
        CONCATENATED label array (bases.py:255-260).
    Returns (embeddings [num_query + n_centroids, d] on the DEVICE, labels, camids) -- the reference moves everything to the
    CPU because its metric runs there (bases.py:262); this engine's metric runs on the GPU."""
    emb = torch.as_tensor(embeddings)
    emb = emb if emb.is_cuda else emb.cuda(non_blocking=True)
    dev = emb.device
    labels = np.asarray(labels)
    camids = np.asarray(camids)
    lab_q, lab_g = labels[:num_query], labels[num_query:]
    n_g = lab_g.shape[0]
    # dense label ids in sorted label order (vectorised host relabelling; everything below is on the device)
    uniq_lab, dense_g = np.unique(lab_g, return_inverse=True)
    n_lab = uniq_lab.shape[0]
    d_lab = torch.from_numpy(dense_g.astype(np.int64)).to(dev, non_blocking=True)
    order = torch.argsort(d_lab, stable=True)                      # gallery rows by (label, row)
    counts = torch.bincount(d_lab, minlength=n_lab)
    seg_start = torch.cumsum(counts, 0) - counts
    if not respect_camids:
        indptr = torch.cat((seg_start, counts.sum()[None]))
        cents = segment_mean_csr(emb, indptr, order + num_query)
        out_lab = np.hstack((lab_q, uniq_lab))
        out_cam = np.hstack((np.zeros_like(lab_q), np.ones_like(out_lab)))
        return torch.cat((emb[:num_query].float(), cents), 0), out_lab, out_cam

    cam_vals, cam_dense = np.unique(camids, return_inverse=True)
    if cam_vals.shape[0] > 63:
        raise NotImplementedError(f"{cam_vals.shape[0]} distinct cameras; camera sets are packed into 64-bit masks")
    cam_dense = cam_dense.astype(np.int64)
    d_cam_g = torch.from_numpy(cam_dense[:n_g]).to(dev, non_blocking=True)          # bases.py:214 quirk
    # queries whose label exists in the gallery -> candidate (label, query camera) pairs, sorted, unique
    pos = np.searchsorted(uniq_lab, lab_q)
    pos = np.clip(pos, 0, max(n_lab - 1, 0))
    has = uniq_lab[pos] == lab_q if n_lab else np.zeros_like(lab_q, dtype=bool)
    d_ql = torch.from_numpy(pos[has].astype(np.int64)).to(dev, non_blocking=True)
    d_qc = torch.from_numpy(cam_dense[:num_query][has]).to(dev, non_blocking=True)
    cand = torch.unique(d_ql * 64 + d_qc)                                            # ascending (label, camera)
    c_lab, c_cam = cand // 64, cand % 64
    present = torch.zeros(n_lab, 64, dtype=torch.bool, device=dev)
    present[d_lab, d_cam_g] = True                                                   # cameras seen per identity
    weights = (torch.ones(64, dtype=torch.int64, device=dev) << torch.arange(64, device=dev))
    mask = (present.to(torch.int64) * weights).sum(1)                                # [n_lab] camera bit sets
    m = mask[c_lab]
    used = m & ~(torch.ones_like(c_cam) << c_cam)
    full = used == m                                                                 # query camera not in the set
    # first occurrence of each distinct set per identity: sets that drop a camera are distinct by construction; the
    # "nothing dropped" set repeats for every query camera outside the identity's cameras -> keep its first one only
    full_idx = torch.nonzero(full).flatten()
    keep = used != 0
    if full_idx.numel():
        fl = c_lab[full_idx]
        first = torch.ones_like(fl, dtype=torch.bool)
        first[1:] = fl[1:] != fl[:-1]
        dup = torch.zeros_like(keep)
        dup[full_idx[~first]] = True
        keep = keep & ~dup
    k_lab, k_used = c_lab[keep], used[keep]
    n_cent = int(k_lab.numel())
    # members: expand every centroid over its identity's segment, keep the rows whose camera is in the set
    seg_len = counts[k_lab]
    cent_of = torch.repeat_interleave(torch.arange(n_cent, device=dev), seg_len)
    offs = torch.arange(cent_of.numel(), device=dev) - torch.repeat_interleave(torch.cumsum(seg_len, 0) - seg_len, seg_len)
    rows = order[seg_start[k_lab][cent_of] + offs]                                   # gallery-relative, ascending per centroid
    sel = ((k_used[cent_of] >> d_cam_g[rows]) & 1).bool()
    members, owner = rows[sel], cent_of[sel]
    mcount = torch.bincount(owner, minlength=n_cent)
    indptr = torch.cat((torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(mcount, 0)))
    cents = segment_mean_csr(emb, indptr, members + num_query)
    out_emb = torch.cat((emb[:num_query].float(), cents), 0)
    k_lab_h, k_used_h = k_lab.cpu().numpy(), k_used.cpu().numpy()
    out_lab = np.hstack((lab_q, uniq_lab[k_lab_h]))
    bits = (k_used_h[:, None] >> np.arange(cam_vals.shape[0])[None, :]) & 1
    cent_cam = [cam_vals[np.nonzero(b)[0]].tolist() for b in bits]
    out_cam = [[c] for c in camids[:num_query].tolist()] + cent_cam
    return out_emb, out_lab, out_cam
