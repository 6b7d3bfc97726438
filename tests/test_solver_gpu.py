"""GPU: fused optimizer step (solver/build.py drop-in) against torch.optim.Adam / SGD."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _C(dict):
    __getattr__ = dict.__getitem__


def _hp(**over):
    s = _C(OPTIMIZER_NAME="Adam", BASE_LR=3.5e-4, WEIGHT_DECAY=5e-4, CENTER_LR=0.5, CENTER_LOSS_WEIGHT=5e-4,
           LR_SCHEDULER_NAME="multistep_lr", LR_STEPS=(2, 4), GAMMA=0.1, USE_WARMUP_LR=True, WARMUP_EPOCHS=3,
           MAX_EPOCHS=6, MIN_LR=1e-6)
    s.update(over)
    return _C(SOLVER=s)


def test_fused_adam_and_center_sgd_match_torch():
    from ctl_b200.solver.build import apply_warmup_lr, build_optimizer, build_scheduler

    g = torch.Generator().manual_seed(0)
    shapes = {"backbone.base.conv1.weight": (64, 3, 7, 7), "backbone.base.bn1.weight": (64,),
              "backbone.base.layer4.2.conv3.weight": (2048, 512, 1, 1), "fc_query.weight": (751, 2048), "odd": (8193,),
              "bn.bias": (2048,), "center_loss.centers": (751, 2048)}
    init = {k: torch.randn(s, generator=g) * 0.1 for k, s in shapes.items()}
    hp = _hp()

    def make(fused):
        ps = {k: torch.nn.Parameter(v.clone().cuda()) for k, v in init.items()}
        ps["bn.bias"].requires_grad_(False)  # frozen in the reference (bases.py:83-84)
        if fused:
            opts = build_optimizer(ps.items(), hp)
        else:
            reg = [p for k, p in ps.items() if p.requires_grad and "center" not in k]
            opts = [torch.optim.Adam(reg, lr=hp.SOLVER.BASE_LR, weight_decay=hp.SOLVER.WEIGHT_DECAY),
                    torch.optim.SGD([ps["center_loss.centers"]], lr=hp.SOLVER.CENTER_LR)]
        return ps, opts, build_scheduler(opts[0], hp)

    (pf, of, sf), (pt, ot, st) = make(True), make(False)
    assert [len(o.param_groups[0]["params"]) for o in of] == [5, 1]
    for epoch in range(5):
        for it in range(2):
            grads = {k: torch.randn(s, generator=g) * (0.05 if "center" not in k else 1e-4) for k, s in shapes.items()}
            for ps, opts in ((pf, of), (pt, ot)):
                for k, p in ps.items():
                    p.grad = grads[k].clone().cuda() if p.requires_grad and not (k == "odd" and it == 1 and epoch == 0) else None
                for o in opts:
                    apply_warmup_lr(o, epoch, hp) if o is opts[0] else None
                opts[0].step()
                ps["center_loss.centers"].grad.data *= 1.0 / hp.SOLVER.CENTER_LOSS_WEIGHT  # train_ctl_model.py:157-158
                opts[1].step()
        sf.step()
        st.step()
        assert of[0].param_groups[0]["lr"] == ot[0].param_groups[0]["lr"]
    torch.cuda.synchronize()
    for k in shapes:
        a, b = pf[k].detach().cpu().numpy(), pt[k].detach().cpu().numpy()
        np.testing.assert_allclose(a, b, rtol=2e-5, atol=2e-7, err_msg=k)
    # optimizer state is torch.optim.Adam's: a reference state_dict loads into the fused optimizer and back
    sd = ot[0].state_dict()
    of[0].load_state_dict(sd)
    assert set(of[0].state_dict()["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
