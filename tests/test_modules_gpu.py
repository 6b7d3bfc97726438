"""GPU: the reference-named module surface (Baseline, CTLModel hooks) on the B200 engine."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ctl_oracle as O
from oracle.make_golden import DIM, LOSS_CASES, NUM_CLASSES, head_state

pytestmark = pytest.mark.gpu


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def _cfg(**over):
    c = _Cfg(
        MODEL=_Cfg(NAME="resnet50", LAST_STRIDE=1, PRETRAINED=False, PRETRAIN_PATH="", BACKBONE_EMB_SIZE=2048,
                   USE_CENTROIDS=False, KEEP_CAMID_CENTROIDS=True, RESUME_TRAINING=False),
        SOLVER=_Cfg(MARGIN=0.5, DISTANCE_FUNC="euclidean", CENTER_LOSS_WEIGHT=5e-4, QUERY_XENT_WEIGHT=1.0,
                    QUERY_CONTRASTIVE_WEIGHT=1.0, CENTROID_CONTRASTIVE_WEIGHT=1.0),
        DATALOADER=_Cfg(NUM_INSTANCE=4), TEST=_Cfg(FEAT_NORM=True, ONLY_TEST=False, VISUALIZE="no"),
        USE_MIXED_PRECISION=True)
    for k, v in over.items():
        a, b = k.split("__")
        c[a][b] = v
    return c


def test_ctl_model_hooks():
    from ctl_b200.modelling.ctl_model import CTLModel

    torch.manual_seed(0)
    model = CTLModel(_cfg(), num_classes=NUM_CLASSES, num_query=24).cuda()
    # state_dict layout of the reference checkpoint
    keys = list(model.state_dict().keys())
    assert "backbone.base.conv1.weight" in keys and "center_loss.centers" in keys and "fc_query.weight" in keys
    assert "bn.running_var" in keys and len([k for k in keys if k.startswith("backbone.base.")]) == 318
    # validation_step == oracle embed_forward on the same weights
    sd = O.make_trunk_state(seed=3)
    model.backbone.base.load_state_dict(sd)
    model.backbone.invalidate()
    with torch.no_grad():
        model.bn.running_mean.normal_(0, 0.1)
        model.bn.running_var.uniform_(0.5, 1.5)
        model.bn.weight.uniform_(0.5, 1.5)
    x = torch.randn(4, 3, 256, 128, generator=torch.Generator().manual_seed(2))
    out = model.validation_step((x.cuda(), torch.arange(4), torch.zeros(4, dtype=torch.long), torch.arange(4)), 0)
    bn_sd = {k: v.detach().cpu() for k, v in model.bn.state_dict().items()}
    with torch.no_grad():
        ref = O.embed_forward(x, sd, bn_sd)
    scale = float(ref.abs().max())
    assert float((out["emb"].cpu() - ref).abs().max()) <= 1e-2 * scale  # fp16 trunk vs fp32 oracle
    model.train()
    assert model.backbone(x.cuda())[1].requires_grad  # train mode: differentiable B200 training engine
    model.eval()
    # training_step tail from prescribed features == the reference's training_step golden
    name = "p8k4_pad"
    g = load_golden(f"loss_{name}.npz")
    P, K, pad, seed, scale_f = LOSS_CASES[name]
    feats, labels, is_real = O.synth_batch(P, K, DIM, NUM_CLASSES, seed, pad, scale_f)
    hs = head_state(seed)
    with torch.no_grad():
        model.center_loss.centers.copy_(hs["centers"])
        model.bn.weight.copy_(hs["bn_weight"])
        model.bn.bias.copy_(hs["bn_bias"])
        model.fc_query.weight.copy_(hs["fc_weight"])
    f = feats.cuda().requires_grad_(True)
    res = model.training_step_from_features(f, labels.cuda(), is_real.cuda())
    res["loss"].backward()
    np.testing.assert_allclose(float(res["loss"]), float(g["total"]), rtol=1e-4)
    np.testing.assert_allclose(f.grad.cpu().numpy(), g["grad_feats"], rtol=1e-4, atol=1e-4 * np.abs(g["grad_feats"]).max())
    assert model.center_loss.centers.grad is not None and model.fc_query.weight.grad is not None


def test_validation_epoch_end_centroid_metric():
    """validation_epoch_end with MODEL.USE_CENTROIDS (both camid modes) vs the reference goldens."""
    from ctl_b200.modelling.ctl_model import CTLModel

    g = load_golden("centroids.npz")
    nq, ng = int(g["num_q"]), int(g["num_g"])
    feats, pids, cams = O.synth_retrieval(nq, ng, int(g["num_ids"]), DIM, 3.0, 11, num_cams=4)
    for keep, tag in ((False, "nocam"), (True, "cam")):
        model = CTLModel(_cfg(MODEL__USE_CENTROIDS=True, MODEL__KEEP_CAMID_CENTROIDS=keep), num_classes=10, num_query=nq)
        outs = [{"emb": feats[i:i + 200].cuda(), "labels": torch.from_numpy(pids[i:i + 200]),
                 "camid": torch.from_numpy(cams[i:i + 200])} for i in range(0, nq + ng, 200)]
        cmc, mAP, topk = model.validation_epoch_end(outs)
        assert np.array_equal(cmc, g[f"{tag}_cmc"])
        np.testing.assert_allclose(mAP, float(g[f"{tag}_mAP"]), rtol=1e-9)
        np.testing.assert_allclose(topk, g[f"{tag}_topk"], rtol=1e-9)
