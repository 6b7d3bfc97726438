"""GPU: the inference drop-ins (inference/inference_utils.py, create_embeddings.py, get_similar.py of the
reference) against the oracle restatement, including the on-disk formats."""
import numpy as np
import pytest
import torch

from oracle import ctl_oracle as O
from test_modules_gpu import _cfg

pytestmark = pytest.mark.gpu


def _model():
    from ctl_b200.modelling.ctl_model import CTLModel

    torch.manual_seed(0)
    model = CTLModel(_cfg(TEST__IMS_PER_BATCH=3), num_classes=16, num_query=4).cuda().eval()
    sd = O.make_trunk_state(seed=5)
    model.backbone.base.load_state_dict(sd)
    model.backbone.invalidate()
    with torch.no_grad():
        model.bn.running_mean.normal_(0, 0.1)
        model.bn.running_var.uniform_(0.5, 1.5)
    return model, sd


def test_run_inference_matches_oracle_embed_and_keeps_order(tmp_path):
    from ctl_b200.inference import inference_utils as IU

    model, sd = _model()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(7, 3, 128, 64, generator=g)
    paths = [f"/data/{i % 3:04d}_c{i}.jpg" for i in range(7)]
    loader = [(x[i:i + 3], [""] * len(paths[i:i + 3]), paths[i:i + 3]) for i in range(0, 7, 3)]  # ragged last batch
    emb, got_paths = IU.run_inference(model, loader, _cfg(TEST__IMS_PER_BATCH=3), print_freq=10, use_cuda=True)
    assert emb.dtype == np.float32 and emb.shape == (7, 2048) and list(got_paths) == paths
    bn_sd = {k: v.detach().cpu() for k, v in model.bn.state_dict().items()}
    with torch.no_grad():
        ref = O.embed_forward(x, sd, bn_sd).numpy()
    assert np.abs(emb - ref).max() <= 1e-2 * np.abs(ref).max()  # fp16 trunk vs the fp32 oracle
    with pytest.raises(RuntimeError):
        IU._inference(model, loader[0], use_cuda=False)
    # centroids + gallery files (create_embeddings.py:96-109)
    index = IU.create_pid_path_index(paths, lambda p: p.split("/")[-1].split("_")[0])
    assert list(index.keys()) == ["0000", "0001", "0002"] and index["0000"] == [0, 3, 6]
    cents, pids = IU.calculate_centroids(emb, index)
    ref_c, ref_p = O.calculate_centroids_by_pid(emb, index)
    np.testing.assert_allclose(cents, ref_c, rtol=1e-6, atol=1e-6)
    assert pids.dtype.kind == "U" and list(pids) == list(ref_p)
    IU.save_gallery(tmp_path / "gal", cents, pids)
    ge, gp = IU.load_gallery(tmp_path / "gal")
    assert ge.dtype == torch.float32 and tuple(ge.shape) == (3, 2048) and list(gp) == list(pids)


@pytest.mark.parametrize("dist,normalize,topk", [("euclidean", False, 10), ("euclidean", True, 5), ("cosine", False, 7),
                                                 ("euclidean", False, 0)])
def test_get_similar_matches_reference_format(tmp_path, dist, normalize, topk):
    from ctl_b200.inference import inference_utils as IU

    feats, _, _ = O.synth_retrieval(24, 300, 20, dim=256, sigma=1.5, seed=4, dyadic=(dist == "euclidean" and not normalize))
    feats = feats.float()
    q, g = feats[:24].numpy(), feats[24:].numpy()
    qpaths = np.array([f"q{i}.jpg" for i in range(24)])
    gpaths = np.array([f"g{i}.jpg" for i in range(300)])
    out = IU.get_similar(q, qpaths, g, gpaths, dist, topk, normalize)
    idx_o, dst_o = O.topk_similar(q, g, topk if topk else 300, dist, normalize)
    assert list(out.keys()) == list(qpaths)
    exact = dist == "euclidean" and not normalize  # dyadic fixture: bit-exact distances and ranks
    for i, qp in enumerate(qpaths):
        r = out[qp]
        assert set(r.keys()) == {"indices", "paths", "distances"}
        if exact:
            assert np.array_equal(r["indices"], idx_o[i])
            assert np.array_equal(r["distances"], dst_o[i])
        else:
            np.testing.assert_allclose(r["distances"], dst_o[i], rtol=0, atol=1e-5)
            # ranks may only differ where the oracle's own distances are within the tolerance
            diff = r["indices"] != idx_o[i]
            assert np.all(np.abs(dst_o[i][diff] - r["distances"][diff]) <= 1e-5)
        assert np.array_equal(r["paths"], gpaths[r["indices"]])
    IU.save_results(tmp_path / "out", out, q, qpaths)
    back = np.load(tmp_path / "out" / "results.npy", allow_pickle=True).item()
    assert np.array_equal(back["q3.jpg"]["indices"], out["q3.jpg"]["indices"])
    assert np.load(tmp_path / "out" / "query_embeddings.npy").shape == (24, 256)
    assert list(np.load(tmp_path / "out" / "query_paths.npy")) == list(qpaths)
