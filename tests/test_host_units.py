"""CPU: host-side pieces added around the hot path (no GPU, no kernels): parameter sampling of the device transforms,
the stem weight packing, the warm-up LR rule, the pid/path index of the inference scripts."""
import numpy as np
import torch

import ctl_b200  # noqa: F401


class _C(dict):
    __getattr__ = dict.__getitem__


def test_sample_params_ranges_and_mock_rows():
    from ctl_b200.datasets.transforms import sample_params

    p = sample_params(500, 64, 32, prob_flip=0.3, pad=5, re_prob=1.0, is_real=np.r_[np.ones(499), 0],
                      rng=np.random.default_rng(1))
    assert p.dtype == np.int32 and p.shape == (500, 8)
    assert set(np.unique(p[:, 0])) <= {0, 1} and abs(p[:, 0].mean() - 0.3) < 0.08
    assert p[:, 1:3].min() >= 0 and p[:, 1:3].max() <= 10
    assert (p[:, 5] > 0).all() and (p[:, 5] < 64).all() and (p[:, 6] < 32).all()      # re_prob = 1: always erased
    assert (p[:, 3] + p[:, 5] <= 64).all() and (p[:, 4] + p[:, 6] <= 32).all()       # rectangle inside the image
    assert p[-1, 7] == 0 and p[:-1, 7].all()
    none = sample_params(50, 64, 32, re_prob=0.0, rng=np.random.default_rng(2))
    assert (none[:, 5] == 0).all()


def test_pack_stem_fused_layout_matches_the_header_definition():
    from ctl_b200.modelling.backbones.engine import pack_stem_fused

    w = torch.randn(64, 3, 7, 7, generator=torch.Generator().manual_seed(0))
    pk = pack_stem_fused(w)
    assert pk.shape == (28, 64, 8) and pk.dtype == torch.float16
    for c in (0, 5, 13, 27):
        for e in range(8):
            r, s, ch = c // 4, 2 * (c % 4) + e // 4, e % 4
            exp = w[:, ch, r, s].half() if (ch < 3 and s < 7) else torch.zeros(64, dtype=torch.float16)
            assert torch.equal(pk[c, :, e], exp), (c, e)


def test_warmup_rule_and_scheduler_names():
    from ctl_b200.solver.build import apply_warmup_lr, build_scheduler

    hp = _C(SOLVER=_C(USE_WARMUP_LR=True, WARMUP_EPOCHS=10, BASE_LR=1e-3, LR_SCHEDULER_NAME="multistep_lr", LR_STEPS=(2, 3),
                      GAMMA=0.1, MAX_EPOCHS=5, MIN_LR=1e-6))
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-3)
    for epoch, exp in ((0, 1e-4), (4, 5e-4), (9, 1e-3)):
        apply_warmup_lr(opt, epoch, hp)
        assert abs(opt.param_groups[0]["lr"] - exp) < 1e-12
    opt.param_groups[0]["lr"] = 0.123
    apply_warmup_lr(opt, 10, hp)  # past the warm-up: untouched (modelling/bases.py:116)
    assert opt.param_groups[0]["lr"] == 0.123
    assert isinstance(build_scheduler(opt, hp), torch.optim.lr_scheduler.MultiStepLR)
    hp.SOLVER.LR_SCHEDULER_NAME = "cosine_annealing"
    assert isinstance(build_scheduler(opt, hp), torch.optim.lr_scheduler.CosineAnnealingLR)
    hp.SOLVER.LR_SCHEDULER_NAME = "nope"
    try:
        build_scheduler(opt, hp)
        raise AssertionError("expected NotImplementedError")
    except NotImplementedError:
        pass


def test_pid_path_index_and_no_cpu_fallback(tmp_path):
    from ctl_b200.inference import inference_utils as IU

    paths = ["a/0002_c1.jpg", "a/0001_c2.jpg", "a/0002_c3.jpg"]
    idx = IU.create_pid_path_index(paths, lambda p: p.split("/")[-1].split("_")[0])
    assert list(idx.items()) == [("0002", [0, 2]), ("0001", [1])]
    IU.save_gallery(tmp_path, np.ones((2, 4), dtype=np.float32), np.array(["x", "y"]))
    emb, pth = IU.load_gallery(tmp_path)
    assert emb.dtype == torch.float32 and list(pth) == ["x", "y"]
    try:
        IU._inference(None, (torch.zeros(1, 3, 8, 8), [""], ["p"]), use_cuda=False)
        raise AssertionError("expected RuntimeError")
    except RuntimeError as e:
        assert "no CPU path" in str(e)
