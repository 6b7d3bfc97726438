"""GPU parity tests of the retrieval path (through the C ABI) against the oracle and the
golden vectors produced by the unmodified reference.

Parity contract (DESIGN.md section "Parity"):
  (a) distances: |d_gpu - d_ref| <= 4e-6 * (|q|^2 + |g|^2) -- fp32-equivalent arithmetic;
      BIT-EXACT on the dyadic-grid fixtures (every product exact in fp32 and in the fp16 split);
  (b) ranks / top-k / CMC / AP are integer-exact functions of the GPU distances under the
      canonical (distance, gallery index) order: streamed results == oracle applied to the
      materialised GPU matrix, at every size;
  (c) hence identical to the reference's own outputs on the exact fixtures, and within 1e-5
      (mAP) on the float fixtures where the reference's own fp32 sgemm has unordered near-ties.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ctl_oracle as O

pytestmark = pytest.mark.gpu

DIM = 2048


@pytest.fixture(scope="module")
def R():
    from ctl_b200 import retrieval

    return retrieval


def _fixture(name):
    g = load_golden(f"retrieval_{name}.npz")
    nq, ng = int(g["num_q"]), int(g["num_g"])
    feats, pids, cams = O.synth_retrieval(nq, ng, int(g["num_ids"]), DIM, float(g["sigma"]), int(g["seed"]),
                                          dyadic=bool(g["dyadic"]))
    return g, nq, ng, feats, pids, cams


def _eval_from_matrix(d, pids, cams, nq, respect=False):
    idx = O.rank_indices(d)
    return idx, O.eval_func(idx, pids[:nq], pids[nq:], cams[:nq], cams[nq:], 50, respect)


@pytest.mark.parametrize("name", ["dyadic", "ties"])
def test_exact_fixtures_bit_exact_vs_reference(R, name):
    g, nq, ng, feats, pids, cams = _fixture(name)
    q, gal = feats[:nq].cuda(), feats[nq:].cuda()
    d = R.dist_matrix(q, gal).cpu().numpy()
    assert np.array_equal(d, g["dist"]), "distance matrix must be bit-identical to the reference's"
    k = g["topk_idx"].shape[1]
    idx, dst = R.topk_similar(q, gal, k)
    assert np.array_equal(idx.cpu().numpy(), g["topk_idx"].astype(np.int64))
    assert np.array_equal(dst.cpu().numpy(), g["topk_dist"])
    res = R.evaluate_streamed(R.build_planes(q), R.build_planes(gal), pids[:nq], pids[nq:], cams[:nq], cams[nq:])
    assert np.array_equal(res.cmc, g["cmc"])
    np.testing.assert_allclose(res.mAP, float(g["mAP"]), rtol=1e-12)
    np.testing.assert_allclose(res.all_topk, g["all_topk"], rtol=1e-12)
    np.testing.assert_allclose(res.single_performance[:, 2].astype(np.float64), g["ap"], rtol=1e-12)


def test_small_float_fixture(R):
    g, nq, ng, feats, pids, cams = _fixture("small")
    q, gal = feats[:nq].cuda(), feats[nq:].cuda()
    d = R.dist_matrix(q, gal).cpu().numpy()
    np.testing.assert_allclose(d, g["dist"], rtol=0, atol=4e-6 * 2)  # unit vectors: |q|^2+|g|^2 = 2
    cd = R.dist_matrix(q, gal, "cosine").cpu().numpy()
    np.testing.assert_allclose(cd, g["cos_dist"], rtol=0, atol=4e-6)
    # (b): streamed == oracle on the GPU matrix, exactly
    idx_o, (cmc_o, map_o, topk_o, single_o) = _eval_from_matrix(d, pids, cams, nq)
    idx, dst = R.topk_similar(q, gal, 100)
    assert np.array_equal(idx.cpu().numpy(), idx_o[:, :100])
    assert np.array_equal(dst.cpu().numpy(), np.take_along_axis(d, idx_o[:, :100], 1))
    res = R.evaluate_streamed(R.build_planes(q), R.build_planes(gal), pids[:nq], pids[nq:], cams[:nq], cams[nq:])
    assert np.array_equal(res.cmc, cmc_o)
    np.testing.assert_allclose(res.mAP, map_o, rtol=1e-12)
    np.testing.assert_allclose(res.single_performance[:, 2].astype(np.float64), single_o[:, 2].astype(np.float64),
                               rtol=1e-12)
    # (c): and close to the reference's own numbers
    np.testing.assert_allclose(res.mAP, float(g["mAP"]), rtol=1e-5)
    assert np.array_equal(res.cmc, g["cmc"])


@pytest.mark.parametrize("nq,ng,dim,k", [(200, 5000, 256, 100), (130, 4097, 512, 7), (1, 9000, 64, 50),
                                         (257, 300, 128, 300), (64, 17, 2048, 17)])
def test_ragged_shapes_self_consistent(R, nq, ng, dim, k):
    """Edge shapes: rows/cols not multiples of the 128x128 tile or the 16-column group, both
    top-k plans (single pass for ng <= 4096, group-min + threshold passes above)."""
    gen = torch.Generator().manual_seed(nq * 7 + ng)
    q = torch.randn(nq, dim, generator=gen).cuda()
    gal = torch.randn(ng, dim, generator=gen).cuda()
    d = R.dist_matrix(q, gal)
    ref = O.get_euclidean(q.cpu(), gal.cpu())
    scale = float((q * q).sum(1).max() + (gal * gal).sum(1).max())
    assert float((d.cpu() - ref).abs().max()) <= 4e-6 * scale
    order = torch.sort(d, dim=1, stable=True)
    idx, dst = R.topk_similar(q, gal, k)
    assert torch.equal(idx, order.indices[:, :k])
    assert torch.equal(dst, order.values[:, :k])


def test_eval_streamed_matches_matrix_path_with_junk_and_invalid_queries(R):
    """Queries without any positive (skipped by the reference, eval_reid.py:63-65), junk rows
    (same pid & same camera) and camera-set junk (respect_camids) on a 5000-row gallery."""
    nq, ng, nid = 300, 5000, 200
    feats, pids, cams = O.synth_retrieval(nq, ng, nid, 512, 2.0, 3, num_cams=3)
    pids[:10] = 10_000 + np.arange(10)  # identities absent from the gallery
    q, gal = feats[:nq].cuda(), feats[nq:].cuda()
    d = R.dist_matrix(q, gal, normalize=True).cpu().numpy()
    _, (cmc_o, map_o, topk_o, single_o) = _eval_from_matrix(d, pids, cams, nq)
    res = R.evaluate_streamed(R.build_planes(q, normalize=True), R.build_planes(gal, normalize=True),
                              pids[:nq], pids[nq:], cams[:nq], cams[nq:])
    assert np.array_equal(res.cmc, cmc_o)
    np.testing.assert_allclose(res.mAP, map_o, rtol=1e-12)
    assert np.array_equal(res.single_performance[:, 0].astype(np.int64), single_o[:, 0].astype(np.int64))
    np.testing.assert_allclose(res.single_performance[:, 2].astype(np.float64), single_o[:, 2].astype(np.float64),
                               rtol=1e-12)
    # respect_camids: gallery rows carry camera SETS
    g_sets = [[int(c), int((c + 1) % 3)] if i % 2 else [int(c)] for i, c in enumerate(cams[nq:])]
    q_sets = [[int(c)] for c in cams[:nq]]
    idx = O.rank_indices(d)
    cmc_s, map_s, _, single_s = O.eval_func(idx, pids[:nq], pids[nq:], q_sets, g_sets, 50, True)
    res2 = R.evaluate_streamed(R.build_planes(q, normalize=True), R.build_planes(gal, normalize=True),
                               pids[:nq], pids[nq:], q_sets, g_sets, 50, True)
    assert np.array_equal(res2.cmc, cmc_s)
    np.testing.assert_allclose(res2.mAP, map_s, rtol=1e-12)


def test_market_shape_against_reference_golden(R):
    """BASELINE config 3: 3368 x 15913 x 2048, top-100 + CMC/mAP vs the reference's outputs."""
    g, nq, ng, feats, pids, cams = _fixture("market")
    q, gal = feats[:nq].cuda(), feats[nq:].cuda()
    qp, gp = R.build_planes(q), R.build_planes(gal)
    idx, dst, ovf = R.topk(qp, gp, 100)
    assert int(ovf.item()) == 0
    idx, dst = idx.cpu().numpy(), dst.cpu().numpy()
    ref_idx = g["topk_idx"].astype(np.int64)
    assert (idx == ref_idx).mean() > 0.999          # near-ties of the reference's own fp32 sgemm may swap
    np.testing.assert_allclose(dst, g["topk_dist"], rtol=0, atol=8e-6)
    # every disagreement is an epsilon-tie in the REFERENCE's distances
    bad = np.nonzero(idx != ref_idx)
    assert np.all(np.abs(dst[bad] - g["topk_dist"][bad]) <= 8e-6)
    res = R.evaluate_streamed(qp, gp, pids[:nq], pids[nq:], cams[:nq], cams[nq:])
    np.testing.assert_allclose(res.mAP, float(g["mAP"]), rtol=1e-5)
    np.testing.assert_allclose(res.cmc, g["cmc"], rtol=0, atol=1.0 / nq + 1e-7)
    assert np.array_equal(res.single_performance[:, 0].astype(np.int32), g["valid_q"])
    # The swaps themselves, counted and classified: a swap can only change CMC / AP of its query if it exchanges a kept
    # positive with a kept non-positive.  Where no such swap exists inside the top-100 of a query, that query's AP over the
    # first 100 ranks is unaffected; and if no query at all has one, the CMC curve must be EQUAL, not merely within 1/Q.
    n_swapped_entries = int((idx != ref_idx).sum())
    rows = np.unique(bad[0])
    gp_, gc_ = pids[nq:], cams[nq:]
    crossing = 0
    for qi in rows:
        ours_pos = (gp_[idx[qi]] == pids[qi]) & ~((gp_[idx[qi]] == pids[qi]) & (gc_[idx[qi]] == cams[qi]))
        ref_pos = (gp_[ref_idx[qi]] == pids[qi]) & ~((gp_[ref_idx[qi]] == pids[qi]) & (gc_[ref_idx[qi]] == cams[qi]))
        if not np.array_equal(ours_pos, ref_pos):
            crossing += 1
    print(f"market shape: {n_swapped_entries} of {idx.size} top-100 entries differ from the reference ({len(rows)} of {nq} "
          f"queries, every one an epsilon-tie <= 8e-6); swaps that cross a kept positive inside the top-100: {crossing} queries")
    if crossing == 0:
        assert np.array_equal(res.cmc[:50], g["cmc"][:50]), "no swap crosses a positive, yet the CMC curves differ"
    # (b) at full size: streamed top-k == stable sort of the materialised GPU matrix
    d = R.dist_matrix(q, gal)
    order = torch.sort(d, dim=1, stable=True)
    assert np.array_equal(idx, order.indices[:, :100].cpu().numpy())
    # size-independent property: the first-hit rank equals the position of the first kept
    # positive in the full ranking for 64 sampled queries
    ranks = res.ranks
    oi = order.indices.cpu().numpy()
    for qi in range(0, nq, 53):
        keep = ~((pids[nq:][oi[qi]] == pids[qi]) & (cams[nq:][oi[qi]] == cams[qi]))
        hits = (pids[nq:][oi[qi]] == pids[qi])[keep]
        pos = np.nonzero(hits)[0] + 1
        n = len(pos)
        assert np.array_equal(ranks[qi, :n], pos)


def test_errors(R):
    q = torch.randn(8, 64).cuda()
    with pytest.raises(ValueError):
        R.build_planes(torch.randn(8, 63).cuda())       # d % 8
    with pytest.raises(RuntimeError):
        R.build_planes(torch.randn(8, 64))              # host tensor: no CPU fallback
    qp = R.build_planes(q)
    idx, dst, ovf = R.topk(qp, qp, 100)                 # k clamps to ng like indices[:, :topk]
    assert idx.shape == (8, 8)


def test_fused_two_pass_topk_and_eval(R):
    """The two-pass composition (what bench.py times) == the separate entry points."""
    for nq, ng, nid, dim in ((300, 5000, 200, 512), (64, 600, 30, 2048)):
        feats, pids, cams = O.synth_retrieval(nq, ng, nid, dim, 2.0, 13, num_cams=3)
        q, gal = feats[:nq].cuda(), feats[nq:].cuda()
        qp, gp = R.build_planes(q), R.build_planes(gal)
        idx, dst, res = R.topk_and_eval(qp, gp, 100, pids[:nq], pids[nq:], cams[:nq], cams[nq:])
        idx2, dst2, _ = R.topk(qp, gp, 100)
        res2 = R.evaluate_streamed(qp, gp, pids[:nq], pids[nq:], cams[:nq], cams[nq:])
        assert torch.equal(idx, idx2) and torch.equal(dst, dst2)
        assert np.array_equal(res.cmc, res2.cmc) and res.mAP == res2.mAP
        assert np.array_equal(res.ranks, res2.ranks)


def test_eval_many_positives_per_query(R):
    """> 32 positives per query: beyond the register-resident thresholds of the count pass (global
    64-bit search fallback), plus exact distance ties between positives (duplicated gallery rows)."""
    nq, ng, nid = 64, 3000, 20
    feats, pids, cams = O.synth_retrieval(nq, ng, nid, 256, 2.5, 17, num_cams=5)
    feats[nq + 100:nq + 140] = feats[nq + 200:nq + 240]      # 40 duplicated gallery rows -> exact ties
    pids[nq + 100:nq + 140] = pids[nq + 200:nq + 240]
    q, gal = feats[:nq].cuda(), feats[nq:].cuda()
    d = R.dist_matrix(q, gal).cpu().numpy()
    _, (cmc_o, map_o, _, single_o) = _eval_from_matrix(d, pids, cams, nq)
    res = R.evaluate_streamed(R.build_planes(q), R.build_planes(gal), pids[:nq], pids[nq:], cams[:nq], cams[nq:])
    assert res.ranks.shape[1] > 32
    assert np.array_equal(res.cmc, cmc_o)
    np.testing.assert_allclose(res.mAP, map_o, rtol=1e-12)
    np.testing.assert_allclose(res.single_performance[:, 2].astype(np.float64), single_o[:, 2].astype(np.float64), rtol=1e-12)


def test_topk_similar_has_no_capability_cliffs():
    """the reference's `argsort[:, :topk]` works for every topk and every input; topk_similar / get_similar must
    too.  (a) a 6000-row gallery with topk = 500 and topk = 1500 -- outside the streamed kernel's plan (k <= ceil(ng/16)
    merged groups) -> the materialised path; (b) 3000 identical gallery rows -- thousands of exact ties at the k-th
    distance overflow the candidate list -> the materialised path; both in the canonical (distance, index) order."""
    from ctl_b200 import retrieval as R
    from ctl_b200.inference.inference_utils import get_similar

    g = torch.Generator().manual_seed(3)
    q = torch.randn(40, 256, generator=g)
    gal = torch.randn(6000, 256, generator=g)
    d = O.get_euclidean(q.double(), gal.double()).numpy()
    for k in (500, 1500):
        idx, dst = R.topk_similar(q.cuda(), gal.cuda(), k)
        dm = R.dist_matrix(q.cuda(), gal.cuda()).cpu().numpy()
        want = np.argsort(dm, axis=1, kind="stable")[:, :k]
        assert np.array_equal(idx.cpu().numpy(), want)
        assert np.array_equal(dst.cpu().numpy(), np.take_along_axis(dm, want, 1))
        assert float(np.abs(np.take_along_axis(d, want, 1) - dst.cpu().numpy()).max()) < 1e-3
    same = torch.randn(1, 128, generator=g).expand(3000, 128).contiguous()
    idx, dst = R.topk_similar(torch.randn(5, 128, generator=g).cuda(), same.cuda(), 100)
    assert np.array_equal(idx.cpu().numpy(), np.tile(np.arange(100), (5, 1)))  # all tied: ascending gallery index
    out = get_similar(q.numpy(), [f"q{i}" for i in range(40)], gal.numpy(), np.asarray([f"g{i}" for i in range(6000)]), topk=700)
    assert out["q3"]["indices"].shape == (700,) and out["q3"]["paths"][0] == f"g{int(out['q3']['indices'][0])}"


def _same_eval(a, b):
    assert np.array_equal(a.cmc, b.cmc) and a.mAP == b.mAP and np.array_equal(a.all_topk, b.all_topk)
    assert np.array_equal(a.ranks, b.ranks)
    assert np.array_equal(a.single_performance, b.single_performance)


@pytest.mark.parametrize("nq,ng,nid,dim,dyadic", [(700, 9000, 150, 512, False), (300, 6000, 60, 256, True),
                                                   (200, 3000, 40, 2048, False)])
def test_tile_lists_and_pid_sorted_planes_are_bit_identical(R, nq, ng, nid, dim, dyadic):
    """Planes stored in identity order (build_planes(order=pid_order(..)), keys written through g_index_map, queries
    un-permuted) with pass 1 restricted to a tile list (the tiles that can hold a positive + a subset of the gallery for the
    threshold) give the SAME indices, distances, ranks, AP, CMC as the full run on the caller's order -- bit for bit.
    (ng = 3000: the small-gallery plan, pass 1 collects only.)  Duplicated gallery rows put exact ties at positives and
    inside the top-k, where the index order decides."""
    feats, pids, cams = O.synth_retrieval(nq, ng, nid, dim, 2.0, 29, num_cams=4, dyadic=dyadic)
    feats[nq + 50:nq + 90] = feats[nq + 500:nq + 540]
    pids[nq + 50:nq + 90] = pids[nq + 500:nq + 540]
    q, gal = feats[:nq].cuda(), feats[nq:].cuda()
    args = (pids[:nq], pids[nq:], cams[:nq], cams[nq:])
    base = R.topk_and_eval(R.build_planes(q), R.build_planes(gal), 50, *args)
    qo, go = R.pid_order(pids[:nq]), R.pid_order(pids[nq:])
    runs = {
        "pid-sorted, tile lists": R.topk_and_eval(R.build_planes(q, order=qo), R.build_planes(gal, order=go), 50, *args),
        "pid-sorted, every tile": R.topk_and_eval(R.build_planes(q, order=qo), R.build_planes(gal, order=go), 50, *args,
                                                  tile_lists=False),
        "gallery sorted only": R.topk_and_eval(R.build_planes(q), R.build_planes(gal, order=go), 50, *args),
        "queries sorted only": R.topk_and_eval(R.build_planes(q, order=qo), R.build_planes(gal), 50, *args),
    }
    for name, (idx, dst, res) in runs.items():
        assert torch.equal(idx, base[0]), name
        assert torch.equal(dst, base[1]), name
        _same_eval(res, base[2])
    ev0 = R.evaluate_streamed(R.build_planes(q), R.build_planes(gal), *args)
    ev1 = R.evaluate_streamed(R.build_planes(q, order=qo), R.build_planes(gal, order=go), *args)
    _same_eval(ev0, base[2])
    _same_eval(ev1, base[2])
    idx_t, dst_t, ovf = R.topk(R.build_planes(q), R.build_planes(gal), 50)          # threshold from a subset of the tiles
    idx_e, dst_e, ovf_e = R.topk(R.build_planes(q), R.build_planes(gal), 50, exact_threshold_pass=True)
    assert int(ovf.item()) == 0 and int(ovf_e.item()) == 0
    assert torch.equal(idx_t, idx_e) and torch.equal(dst_t, dst_e) and torch.equal(idx_t, base[0])


def test_tile_list_pass_against_the_full_pass(R):
    """Kernel-level: ctl_dist_worklist keeps exactly the tiles whose identity ranges intersect plus every stride-th gallery
    tile (checked against numpy); a pass over that list writes the SAME group minima on the kept
    tiles (the rest stay +inf), collects the same positives, and its tau is an upper bound of the full pass's tau."""
    import ctypes as C

    from ctl_b200 import _native as N

    nq, ng, nid, dim, k = 640, 8192, 64, 1024, 50
    feats, pids, cams = O.synth_retrieval(nq, ng, nid, dim, 2.0, 41, num_cams=4)
    qo, go = R.pid_order(pids[:nq]), R.pid_order(pids[nq:])
    qp = R.build_planes(feats[:nq].cuda(), order=qo)
    gp = R.build_planes(feats[nq:].cuda(), order=go)
    ids = R.encode_ids(pids[:nq], pids[nq:], cams[:nq], cams[nq:], False, "cuda", q_order=qo, g_order=go)
    L = N.lib()
    n_groups, m_tiles, n_tiles = (ng + 15) // 16, (nq + 127) // 128, (ng + 127) // 128
    stride = L.ctl_dist_subset_stride(ng, k)
    assert 1 < stride <= n_tiles // 8
    work = R._tile_list(qp, gp, ids, stride)
    torch.cuda.synchronize()
    w = work.cpu().numpy()
    # expectation in numpy: ranges of the sorted identity arrays, ids ascending (gallery tile major, query tiles fastest)
    qpid, gpid = ids.q_pid.cpu().numpy(), ids.g_pid.cpu().numpy()
    qr = np.array([[qpid[i * 128:(i + 1) * 128].min(), qpid[i * 128:(i + 1) * 128].max()] for i in range(m_tiles)])
    gr = np.array([[gpid[i * 128:(i + 1) * 128].min(), gpid[i * 128:(i + 1) * 128].max()] for i in range(n_tiles)])
    keep = ~((gr[None, :, 1] < qr[:, None, 0]) | (gr[None, :, 0] > qr[:, None, 1])) | (np.arange(n_tiles)[None, :] % stride == 0)
    expect = [nt * m_tiles + mt for nt in range(n_tiles) for mt in range(m_tiles) if keep[mt, nt]]  # id = nt * m_tiles + mt
    assert int(w[0]) == len(expect) and np.array_equal(w[1:1 + len(expect)], np.asarray(expect))
    assert 0.1 < len(expect) / (m_tiles * n_tiles) < 0.6
    out = {}
    for use_list in (False, True):
        gmin = torch.full((nq, n_groups), float("inf"), device="cuda")
        pos = torch.zeros(nq, ids.max_pos, dtype=torch.int64, device="cuda")
        cnt = torch.zeros(nq, dtype=torch.int32, device="cuda")
        ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
        p = N.PassDesc(gmin=gmin.data_ptr(), pos_keys=pos.data_ptr(), pos_count=cnt.data_ptr(), q_pid=ids.q_pid.data_ptr(),
                       q_cam=ids.q_cam.data_ptr(), g_pid=ids.g_pid.data_ptr(), g_cammask=ids.g_mask.data_ptr(),
                       max_pos=ids.max_pos, overflow=ovf.data_ptr(), tile_list=work.data_ptr() if use_list else None,
                       g_index_map=gp.order.data_ptr())
        N.check(L.ctl_dist_pass(qp.ptr, nq, gp.ptr, ng, dim, qp.flags, C.byref(p), N.stream_ptr()))
        N.check(L.ctl_sort_key_rows(pos.data_ptr(), cnt.data_ptr(), nq, ids.max_pos, N.stream_ptr()))
        tau = torch.empty(nq, device="cuda")
        N.check(L.ctl_select_tau(gmin.data_ptr(), nq, n_groups, 1, k, tau.data_ptr(), N.stream_ptr()))
        torch.cuda.synchronize()
        assert int(ovf.item()) == 0
        out[use_list] = (gmin.cpu().numpy(), pos.cpu().numpy(), cnt.cpu().numpy(), tau.cpu().numpy())
    (g0, p0, c0, t0), (g1, p1, c1, t1) = out[False], out[True]
    assert np.array_equal(c0, c1)
    col = np.arange(p0.shape[1])[None, :]
    assert np.array_equal(np.where(col < c0[:, None], p0, 0), np.where(col < c1[:, None], p1, 0))
    kept = np.repeat(np.repeat(keep, 128, 0)[:nq], 8, 1)[:, :n_groups]   # 8 groups of 16 columns per gallery tile
    assert np.array_equal(g1[kept], g0[kept]) and np.isinf(g1[~kept]).all() and np.isfinite(g0).all()
    assert (t1 >= t0).all() and np.isfinite(t1).all()
    # a tile list is refused where every tile is needed
    bad = N.PassDesc(dist_out=gmin.data_ptr(), ld_out=ng, tile_list=work.data_ptr())
    assert L.ctl_dist_pass(qp.ptr, nq, gp.ptr, ng, dim, qp.flags, C.byref(bad), N.stream_ptr()) == -1


def test_topk_eval_session_replays_bit_identical(R):
    """TopkEvalSession = the launch sequence of topk_and_eval captured once in a CUDA graph: successive query sets against
    the resident gallery give exactly topk_and_eval's indices, distances, ranks, AP and CMC."""
    nq, ng, nid, dim = 500, 7000, 120, 512
    feats, pids, cams = O.synth_retrieval(nq, ng, nid, dim, 2.0, 53, num_cams=4)
    gal = feats[nq:].cuda()
    args = (pids[:nq], pids[nq:], cams[:nq], cams[nq:])
    sess = R.TopkEvalSession(gal, nq, 50, *args)
    g = torch.Generator().manual_seed(7)
    for rep in range(3):
        q = torch.nn.functional.normalize(feats[:nq] + 0.3 * rep * torch.randn(nq, dim, generator=g), dim=1).cuda()
        idx, dst, res = sess(q)
        idx0, dst0, res0 = R.topk_and_eval(R.build_planes(q), R.build_planes(gal), 50, *args)
        assert torch.equal(idx, idx0) and torch.equal(dst, dst0)
        _same_eval(res, res0)
