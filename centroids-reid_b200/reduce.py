"""Per-identity centroid builders on the segmented-mean kernel (ctl_segment_mean).

Host-side mirrors of modelling/bases.py:92-95,179-262 and
inference/inference_utils.py:147-159.  The grouping logic (which rows form a centroid) is the
reference's own host logic, kept verbatim in behaviour including its quirks; only the
reductions run on the device, as ONE launch over a CSR description of all groups.
"""
from __future__ import annotations

from collections import defaultdict

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


def validation_create_centroids(embeddings, labels, camids, num_query, respect_camids=False):
    """ModelBase.validation_create_centroids (modelling/bases.py:179-262).

    Returns (embeddings [num_query + n_centroids, d] on the DEVICE, labels, camids) -- the
    reference moves everything to the CPU here because its metric runs there (bases.py:262);
    this engine's metric runs on the GPU, so the features stay resident.
    """
    emb = torch.as_tensor(embeddings)
    emb = emb if emb.is_cuda else emb.cuda(non_blocking=True)
    labels = np.asarray(labels)
    camids = np.asarray(camids)
    lab_q, lab_g = labels[:num_query], labels[num_query:]
    l2i, l2i_q = defaultdict(list), defaultdict(list)
    for i, l in enumerate(lab_g.tolist()):
        l2i[l].append(i)
    for i, l in enumerate(lab_q.tolist()):
        l2i_q[l].append(i)
    groups, cent_lab, cent_cam = [], [], []
    for label in sorted(l2i.keys()):
        inds = np.asarray(l2i[label])
        if respect_camids:
            seen = set()
            cam_g = camids[inds]  # reference quirk (bases.py:214): FULL camid array, gallery-relative indices
            cam_q = camids[l2i_q[label]]
            for cur in sorted(np.unique(cam_q).tolist()):
                sel = np.where(cam_g != cur)[0]
                if sel.shape[0] == 0:
                    continue
                used = tuple(sorted(np.unique([c for c in cam_g.tolist() if c != cur]).tolist()))
                if used not in seen:
                    seen.add(used)
                    groups.append(num_query + inds[sel])
                    cent_cam.append(list(used))
                    cent_lab.append(label)
        else:
            cent_lab.append(label)
            groups.append(num_query + inds)
    cents = segment_mean(emb, groups)
    out_emb = torch.cat((emb[:num_query].float(), cents), 0)
    out_lab = np.hstack((lab_q, np.asarray(cent_lab)))
    if respect_camids:
        out_cam = [[c] for c in camids[:num_query].tolist()] + cent_cam
    else:
        # bases.py:255-260 sizes the dummy gallery camids from the concatenated label array
        out_cam = np.hstack((np.zeros_like(lab_q), np.ones_like(out_lab)))
    return out_emb, out_lab, out_cam
