"""Oracle: parameter-update half of `image.optimize_parameters` (TEST INFRASTRUCTURE).

Restates, as explicit per-tensor arithmetic in the *same operation order* as the
reference's foreach passes (so fp32 rounding matches bit for bit on CPU):
  * torch.nn.utils.clip_grad_norm_(params, 1.0)         image.py:540-544
  * adan_sf.step / _multi_tensor_adan                    optimizers/adan_sf.py:138-330
  * AveragedModel + get_ema_multi_avg_fn(decay)          image.py:80-87, 661-662
"""
from __future__ import annotations

import math

import torch
from torch import Tensor


def clip_grad_norm(grads: list, max_norm: float = 1.0) -> Tensor:
    """torch.nn.utils.clip_grad_norm_ semantics (L2, error_if_nonfinite=False):
    coef = clamp(max_norm / (||g|| + 1e-6), max=1); grads *= coef.  Returns ||g||."""
    norms = [g.norm(2) for g in grads]
    total = torch.stack(norms).norm(2)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


class AdanSFState:
    """Param-group scalars + per-parameter state of adan_sf (adan_sf.py:76-89, 217-224)."""

    def __init__(self, params: list, lr=1.6e-3, betas=(0.98, 0.92, 0.99), eps=1e-8, weight_decay=0.02,
                 max_grad_norm=0.0, warmup_steps=0, r=0.0, weight_lr_power=2.0, schedule_free=True):
        self.params = params
        self.lr, self.betas, self.eps, self.weight_decay = lr, tuple(betas), eps, weight_decay
        self.max_grad_norm, self.warmup_steps, self.r = max_grad_norm, warmup_steps, r
        self.weight_lr_power, self.schedule_free = weight_lr_power, schedule_free
        self.step = 0
        self.weight_sum = 0.0
        self.lr_max = -1.0
        self.exp_avg = [torch.zeros_like(p) for p in params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in params]
        self.exp_avg_diff = [torch.zeros_like(p) for p in params]
        self.z = None
        self.neg_pre_grad = None


def adan_sf_scalars(st: AdanSFState) -> dict:
    """Host-side scalar schedule of adan_sf.step (adan_sf.py:176-211). Advances st.step."""
    b1, b2, b3 = st.betas
    st.step += 1
    bc1 = 1.0 - b1 ** st.step
    bc2 = 1.0 - b2 ** st.step
    bc3 = 1.0 - b3 ** st.step
    if st.schedule_free:
        sched = st.step / st.warmup_steps if st.step < st.warmup_steps else 1.0
        lr = st.lr * sched * math.sqrt(bc3)
        st.lr_max = max(lr, st.lr_max)
        weight = (st.step ** st.r) * (st.lr_max ** st.weight_lr_power)
        st.weight_sum = st.weight_sum + weight
        try:
            ckp1 = weight / st.weight_sum
        except ZeroDivisionError:
            ckp1 = 0
    else:
        ckp1 = None
    return {"bc1": bc1, "bc2": bc2, "bc3_sqrt": math.sqrt(bc3), "ckp1": ckp1}


@torch.no_grad()
def adan_sf_step(st: AdanSFState, grads: list) -> None:
    """adan_sf.step + _multi_tensor_adan (adan_sf.py:138-330), max_grad_norm handled as
    in 142-163 (global-norm clip *inside* the optimizer, off by default)."""
    if st.max_grad_norm > 0:
        gn = torch.zeros(1)
        for g in grads:
            gn.add_(g.pow(2).sum())
        gn = torch.sqrt(gn)
        clip = torch.clamp(torch.tensor(st.max_grad_norm) / (gn + st.eps), max=1.0).item()
    else:
        clip = 1.0
    sc = adan_sf_scalars(st)
    b1, b2, b3 = st.betas
    if st.z is None:
        st.z = [p.clone() for p in st.params]
    if st.neg_pre_grad is None or st.step == 1:
        st.neg_pre_grad = [g.clone().mul_(-clip) for g in grads]
    lr, wd, eps, ckp1 = st.lr, st.weight_decay, st.eps, sc["ckp1"]
    for p, g, m, n, d, z, npg in zip(st.params, grads, st.exp_avg, st.exp_avg_sq, st.exp_avg_diff,
                                     st.z, st.neg_pre_grad):
        g.mul_(clip)
        npg.add_(g)
        m.mul_(b1).add_(g, alpha=1 - b1)
        d.mul_(b2).add_(npg, alpha=1 - b2)
        npg.mul_(b2).add_(g)
        n.mul_(b3).addcmul_(npg, npg, value=1 - b3)
        denom = n.sqrt().div_(sc["bc3_sqrt"]).add_(eps)
        p.mul_(1 - lr * wd)
        if st.schedule_free:
            step_size_diff = lr * (b2 / sc["bc2"] * (1 - ckp1))
            step_size = lr * (sc["bc1"] * (1 - ckp1))
            p.lerp_(z, weight=ckp1)
            p.addcdiv_(m, denom, value=-step_size)
            p.addcdiv_(d, denom, value=-step_size_diff)
            z.sub_(g, alpha=lr)
        else:
            step_size_diff = lr * b2 / sc["bc2"]
            step_size = lr / sc["bc1"]
            p.addcdiv_(m, denom, value=-step_size)
            p.addcdiv_(d, denom, value=-step_size_diff)
        npg.zero_().add_(g, alpha=-1.0)


@torch.no_grad()
def adan_sf_eval(st: AdanSFState) -> None:
    """adan_sf.eval, adan_sf.py:112-123: p <- lerp(p, z, 1 - 1/beta1)."""
    if st.z is not None:
        for p, z in zip(st.params, st.z):
            p.lerp_(z, weight=1 - 1 / st.betas[0])


@torch.no_grad()
def adan_sf_train(st: AdanSFState) -> None:
    """adan_sf.train, adan_sf.py:125-136: p <- lerp(p, z, 1 - beta1)."""
    if st.z is not None:
        for p, z in zip(st.params, st.z):
            p.lerp_(z, weight=1 - st.betas[0])


class AdamWState:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.params, self.lr, self.betas, self.eps, self.weight_decay = params, lr, tuple(betas), eps, weight_decay
        self.step = 0
        self.exp_avg = [torch.zeros_like(p) for p in params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in params]


@torch.no_grad()
def adamw_step(st: AdamWState, grads: list) -> None:
    """torch.optim.AdamW single-tensor update (base.py:154-155 selects it for C5)."""
    st.step += 1
    b1, b2 = st.betas
    bc1 = 1 - b1 ** st.step
    bc2 = 1 - b2 ** st.step
    for p, g, m, v in zip(st.params, grads, st.exp_avg, st.exp_avg_sq):
        p.mul_(1 - st.lr * st.weight_decay)
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(st.eps)
        p.addcdiv_(m, denom, value=-st.lr / bc1)


class EMAState:
    """AveragedModel(multi_avg_fn=get_ema_multi_avg_fn(decay)) over *parameters only*
    (use_buffers=False): first update copies, later ones lerp by (1 - decay)."""

    def __init__(self, params: list, decay: float = 0.999):
        self.decay = decay
        self.n_averaged = 0
        self.avg = [p.detach().clone() for p in params]

    @torch.no_grad()
    def update(self, params: list) -> None:
        if self.n_averaged == 0:
            for a, p in zip(self.avg, params):
                a.copy_(p)
        else:
            for a, p in zip(self.avg, params):
                a.lerp_(p, 1 - self.decay)
        self.n_averaged += 1
