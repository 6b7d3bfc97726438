"""Drop-in for solver/build.py: `build_optimizer(named_parameters, hparams)` -> [Adam over everything but the
centers, SGD over the centers], `build_scheduler`, plus the warm-up rule of ModelBase.optimizer_step
(modelling/bases.py:102-133).

The Adam step over the ~160 parameter tensors is ONE multi-tensor kernel launch (`ctl_adam_multi_step`) instead of
torch's per-tensor op chains; state keys (`step`, `exp_avg`, `exp_avg_sq`) and `param_groups` are those of
torch.optim.Adam, so optimizer state_dicts of the reference load unchanged and torch's LR schedulers drive it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _native as N

_CHUNK = 8192  # CTL_OPT_CHUNK


def _upload_table(rows, device) -> torch.Tensor:
    """int64 descriptor table -> device WITHOUT a stream synchronisation: a pageable `.to(device)` is a blocking copy that
    waits for everything already enqueued on the stream (it cost the training step 2.6 ms); a pinned source with
    non_blocking=True is an ordinary asynchronous copy."""
    host = torch.from_numpy(np.asarray(rows, dtype=np.int64)).pin_memory()
    dev = host.to(device, non_blocking=True)
    dev._ctl_host = host  # keep the pinned source alive until the copy has certainly run
    return dev


def _bump_version(tensors):
    """The kernels write parameters through raw pointers; tell torch (autograd's saved-tensor checks, and the eval
    engine's version-keyed weight cache in modelling/baseline.py) that they changed."""
    inc = getattr(torch.autograd.graph, "increment_version", None)
    for t in tensors:
        if inc is not None:
            inc(t)
        else:  # pragma: no cover - very old torch
            t.add_(0)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) semantics (L2 weight decay, no amsgrad)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_mul = 1.0  # e.g. 1 / loss_scale
        self.skip_flag = None  # device int32[1] (DynamicLossScaler): non-zero at execution time -> the kernels do nothing

    def undo_step_count(self):
        """A step that the device skipped (overflow) must not count for the bias correction: DynamicLossScaler calls this
        when it learns, one step late, that the previous step was skipped."""
        for st in self.state.values():
            if "step" in st and float(st["step"]) > 0:
                st["step"] -= 1

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = N.lib()
        for group in self.param_groups:
            by_step, keep = {}, []
            for p in group["params"]:
                if p.grad is None:
                    continue
                N.require_cuda(p)
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError("FusedAdam expects contiguous fp32 parameters and gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                keep.append(g)
                # tensors that skipped steps (no gradient) carry their own bias correction: one launch per step count
                by_step.setdefault(int(st["step"]), []).append(
                    [p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(), 0])
            b1, b2 = group["betas"]
            for step, rows in by_step.items():
                chunks = 0
                for r in rows:
                    r[5] = chunks
                    chunks += (r[4] + _CHUNK - 1) // _CHUNK
                table = _upload_table(rows, group["params"][0].device)
                N.check(L.ctl_adam_multi_step(table.data_ptr(), len(rows), chunks, float(group["lr"]), float(b1), float(b2),
                                              float(group["eps"]), float(group["weight_decay"]), step, float(self.grad_mul),
                                              N.ptr(self.skip_flag), N.stream_ptr()))
                keep.append(table)
            _bump_version([p for p in group["params"] if p.grad is not None])
        return loss


class CenterSGD(torch.optim.Optimizer):
    """torch.optim.SGD(params, lr) without momentum, with the reference's gradient rescale folded in:
    `param.grad *= 1 / CENTER_LOSS_WEIGHT; opt_center.step()` (train_ctl_model.py:157-159) == step(grad_mul=1/w)."""

    def __init__(self, params, lr, grad_mul: float = 1.0):
        super().__init__(params, dict(lr=lr))
        self.grad_mul = grad_mul
        self.skip_flag = None  # see FusedAdam.skip_flag

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                N.require_cuda(p)
                g = p.grad.contiguous()
                N.check(N.lib().ctl_sgd_step(p.data_ptr(), g.data_ptr(), p.numel(), float(group["lr"]), float(self.grad_mul),
                                             N.ptr(self.skip_flag), N.stream_ptr()))
                _bump_version([p])
        return loss


class DynamicLossScaler:
    """torch.cuda.amp.GradScaler semantics for the fp16 trunk backward (the reference trains under PL native AMP,
    utils/misc.py:111): the trunk's data / weight gradients are computed on `scale * dLoss/dfeat`, un-scaled in fp32, and
    a step whose gradients contain inf / NaN is SKIPPED and halves the scale; `growth_interval` clean steps double it.
    Defaults are GradScaler's (init 2^16, x2 / x0.5, interval 2000).

    Unlike GradScaler.step (which reads found_inf back every step), nothing here synchronises: the scale, the growth
    tracker and the overflow flag live on the device, the optimizer kernels skip themselves when the flag is set
    (`skip_flag`), `ctl_loss_scale_update` is GradScaler.update() as a one-thread kernel, and the only thing the host
    needs -- Adam's step COUNT must not include skipped steps -- is corrected one step late from a pinned copy of the
    flag (`settle`), before the next optimizer launch and after a wait on an event that completed a whole step ago."""

    def __init__(self, device, init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000,
                 base_scale=1024.0, enabled=True):
        self.device = torch.device(device)
        self.base_scale = float(base_scale)  # the fixed scale baked into the training engine's kernels / CUDA graphs
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, int(growth_interval)
        self.enabled = enabled
        self.state = torch.tensor([init_scale, init_scale / base_scale, base_scale / init_scale], dtype=torch.float32,
                                  device=self.device)  # scale, ratio, 1 / ratio
        self._ints = torch.zeros(3, dtype=torch.int32, device=self.device)  # tracker, found_inf, last_found
        self._pinned = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._event = None
        self._tables = {}
        self.skipped_steps = 0

    # device scalars consumed by the backward (no host value is ever baked into a launch)
    @property
    def ratio(self) -> torch.Tensor:
        return self.state[1]

    @property
    def inv_ratio(self) -> torch.Tensor:
        return self.state[2]

    @property
    def flag(self) -> torch.Tensor:
        return self._ints[1:2]

    @property
    def scale(self) -> float:
        return float(self.state[0].item())  # synchronises: logging / tests / state_dict only

    def _table(self, grads):
        key = tuple((g.data_ptr(), g.numel()) for g in grads)
        t = self._tables.get(key)
        if t is None:
            rows, chunks = [], 0
            for g in grads:
                rows.append([g.data_ptr(), g.numel(), chunks])
                chunks += (g.numel() + _CHUNK - 1) // _CHUNK
            t = (_upload_table(rows, self.device), len(rows), chunks)
            if len(self._tables) > 8:
                self._tables.clear()
            self._tables[key] = t
        return t

    def check(self, grads, mul: float = 1.0, mul_dev=None):
        """OR-accumulates the overflow flag over `grads` (fp32, contiguous) in ONE multi-tensor launch; a factor
        mul * (*mul_dev) != 1 rescales them in place (mul_dev: a device scalar, e.g. `inv_ratio`)."""
        grads = [g for g in grads if g is not None]
        if not grads:
            return
        for g in grads:
            if g.dtype != torch.float32 or not g.is_contiguous():
                raise TypeError("DynamicLossScaler.check expects contiguous fp32 gradients")
        table, n, chunks = self._table(grads)
        N.check(N.lib().ctl_grad_check_multi(table.data_ptr(), n, chunks, float(mul), N.ptr(mul_dev), self.flag.data_ptr(),
                                             N.stream_ptr()))

    def settle(self, *optimizers) -> bool:
        """Call before launching an optimizer step: if the PREVIOUS step was skipped on the device, take it back out of
        the optimizers' step counters.  Waits on an event recorded a whole step ago (no pipeline bubble)."""
        if self._event is None:
            return False
        self._event.synchronize()
        self._event = None
        skipped = bool(int(self._pinned[0]))
        if skipped:
            self.skipped_steps += 1
            for o in optimizers:
                if hasattr(o, "undo_step_count"):
                    o.undo_step_count()
        return skipped

    def update(self):
        """GradScaler.update() on the device (after the optimizer launches of this step), plus the asynchronous copy of
        this step's verdict that `settle` reads before the next one."""
        i = self._ints
        N.check(N.lib().ctl_loss_scale_update(self.state.data_ptr(), i[0:1].data_ptr(), i[1:2].data_ptr(), i[2:3].data_ptr(),
                                              self.base_scale, float(self.growth_factor), float(self.backoff_factor),
                                              self.growth_interval, N.stream_ptr()))
        self._pinned.copy_(i[2:3], non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()

    def state_dict(self):
        return {"scale": self.scale, "growth_factor": self.growth_factor, "backoff_factor": self.backoff_factor,
                "growth_interval": self.growth_interval, "_growth_tracker": int(self._ints[0].item())}

    def load_state_dict(self, sd):
        sc = float(sd["scale"])
        self.state.copy_(torch.tensor([sc, sc / self.base_scale, self.base_scale / sc]))
        self._ints[0] = int(sd.get("_growth_tracker", 0))


def build_optimizer(named_parameters, hparams, fold_center_rescale: bool = False):
    """solver/build.py:9-47.  Returns [model_optimizer, optimizer_center].  With fold_center_rescale=True the center
    optimizer applies the 1 / CENTER_LOSS_WEIGHT gradient rescale itself (then the caller must NOT rescale)."""
    regular, regular_names, center, center_names = [], [], [], []
    for name, parameter in named_parameters:
        if parameter.requires_grad is False:
            print(f"Parameter {name} does not need a Grad. Excluding from the optimizer...")
            continue
        if "center" in name:
            center.append(parameter)
            center_names.append(name)
        else:
            regular.append(parameter)
            regular_names.append(name)
    if hparams.SOLVER.OPTIMIZER_NAME != "Adam":
        raise NotImplementedError(f"No such optimizer {hparams.SOLVER.OPTIMIZER_NAME}")
    model_optimizer = FusedAdam([{"params": regular, "names": regular_names}], lr=hparams.SOLVER.BASE_LR,
                                weight_decay=hparams.SOLVER.WEIGHT_DECAY)
    mul = 1.0 / hparams.SOLVER.CENTER_LOSS_WEIGHT if fold_center_rescale else 1.0
    optimizer_center = CenterSGD([{"params": center, "names": center_names}], lr=hparams.SOLVER.CENTER_LR, grad_mul=mul)
    return [model_optimizer, optimizer_center]


def build_scheduler(model_optimizer, hparams):
    """solver/build.py:50-63 (torch's schedulers drive `param_groups[...]['lr']` of the fused optimizer)."""
    if hparams.SOLVER.LR_SCHEDULER_NAME == "cosine_annealing":
        return torch.optim.lr_scheduler.CosineAnnealingLR(model_optimizer, hparams.SOLVER.MAX_EPOCHS,
                                                          eta_min=hparams.SOLVER.MIN_LR)
    if hparams.SOLVER.LR_SCHEDULER_NAME == "multistep_lr":
        return torch.optim.lr_scheduler.MultiStepLR(model_optimizer, milestones=hparams.SOLVER.LR_STEPS,
                                                    gamma=hparams.SOLVER.GAMMA)
    raise NotImplementedError(f"No such scheduler {hparams.SOLVER.LR_SCHEDULER_NAME}")


def apply_warmup_lr(optimizer, epoch: int, hparams) -> None:
    """ModelBase.optimizer_step's rule (modelling/bases.py:115-121): lr = min(1, (epoch + 1) / WARMUP_EPOCHS) * BASE_LR
    during the warm-up epochs (applied to whichever optimizer is stepped, like the reference)."""
    if hparams.SOLVER.USE_WARMUP_LR and epoch < hparams.SOLVER.WARMUP_EPOCHS:
        lr_scale = min(1.0, float(epoch + 1) / float(hparams.SOLVER.WARMUP_EPOCHS))
        for pg in optimizer.param_groups:
            pg["lr"] = lr_scale * hparams.SOLVER.BASE_LR
