"""Schedule-free Adan on ONE fused multi-tensor kernel — drop-in for
neosr/optimizers/adan_sf.py (same constructor, param-group keys, per-parameter state names
`exp_avg, exp_avg_sq, exp_avg_diff, z, neg_pre_grad`, `.train()/.eval()`), so `*.state`
checkpoints interchange (SURVEY.md §8b "Optimizer").

`step()` additionally accepts `clip_max_norm` and `ema=(ema_params, decay, first)` so the
model can fuse `clip_grad_norm_` (image.py:540-544) and the EMA update (image.py:661-662)
into the same pass: the reference's ~57 foreach tensor passes become one read+write of each
state tensor."""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch.optim.optimizer import Optimizer

from .. import _lib
from .._lib import NsrAdanSF
from .. import ops as ops_mod
from ..ops import _stream
from ._table import ParamTable, grad_sumsq


class adan_sf(Optimizer):
    def __init__(self, params, lr: float = 1.6e-3, betas=(0.98, 0.92, 0.99), eps: float = 1e-8,
                 weight_decay: float = 0.02, max_grad_norm: float = 0.0, warmup_steps: int = 0, r: float = 0.0,
                 weight_lr_power: float = 2.0, schedule_free: bool = True, **kwargs) -> None:
        if not max_grad_norm >= 0.0:
            raise ValueError(f"Invalid Max grad norm: {max_grad_norm}")
        if not lr >= 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not eps >= 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        for i in range(3):
            if not 0.0 <= betas[i] < 1.0:
                raise ValueError(f"Invalid beta parameter at index {i}: {betas[i]}")
        if max_grad_norm > 0:
            raise NotImplementedError("neosr_b200.adan_sf: the optimizer-internal max_grad_norm clip is not built; "
                                      "use the model's grad_clip (clip_grad_norm_ 1.0), which is fused")
        defaults = {"lr": lr, "betas": betas, "eps": eps, "r": r, "weight_decay": weight_decay,
                    "max_grad_norm": max_grad_norm, "warmup_steps": warmup_steps, "train_mode": True,
                    "weight_sum": 0.0, "lr_max": -1.0, "weight_lr_power": weight_lr_power,
                    "schedule_free": schedule_free}
        super().__init__(params, defaults)
        self._tables: dict = {}
        self._sumsq = None
        self._hp_slots: dict = {}
        self._plan: list = []
        self._fresh: list = []
        self._clip = 0.0

    def __setstate__(self, state) -> None:
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault("schedule_free", True)

    @torch.no_grad()
    def eval(self) -> None:  # adan_sf.py:112-123
        for group in self.param_groups:
            beta1 = group["betas"][0]
            if group["train_mode"]:
                for p in group["params"]:
                    st = self.state[p]
                    if "z" in st:
                        p.lerp_(end=st["z"], weight=1 - 1 / beta1)
                group["train_mode"] = False

    @torch.no_grad()
    def train(self) -> None:  # adan_sf.py:125-136
        for group in self.param_groups:
            beta1 = group["betas"][0]
            if not group["train_mode"]:
                for p in group["params"]:
                    st = self.state[p]
                    if "z" in st:
                        p.lerp_(end=st["z"], weight=1 - beta1)
                group["train_mode"] = True

    def _all_grads_table(self) -> ParamTable:
        rows = [{"p": p.detach(), "g": p.grad} for g in self.param_groups for p in g["params"] if p.grad is not None]
        t = self._tables.setdefault("norm", ParamTable())
        return t.build(rows)

    # ---------------------------------------------------------------------------------------
    # step() = prepare() (host: advance the schedule, derive the scalars exactly where the reference
    # does, adan_sf.py:176-211/289-330) + launch() (device: grad-norm + ONE fused kernel).  In
    # CUDA-graph mode the model calls prepare(..., to_device=True) every iteration (a 72-byte async
    # copy) and replays a graph that contains launch().
    @torch.no_grad()
    def prepare(self, *, clip_max_norm: float | None = None, ema=None, to_device: bool = False) -> None:
        ema_iter = iter(ema[0]) if ema is not None else None
        self._plan = []
        self._fresh = []
        self._clip = float(clip_max_norm or 0.0)
        for gi, group in enumerate(self.param_groups):
            beta1, beta2, beta3 = group["betas"]
            group["step"] = group["step"] + 1 if "step" in group else 1
            step = group["step"]
            bc1 = 1.0 - beta1 ** step
            bc2 = 1.0 - beta2 ** step
            bc3 = 1.0 - beta3 ** step
            lr = group["lr"]
            if self.defaults["schedule_free"]:
                warm = group["warmup_steps"]
                sched = step / warm if step < warm else 1.0
                lr_s = group["lr"] * sched * math.sqrt(bc3)
                lr_max = group["lr_max"] = max(lr_s, group["lr_max"])
                weight = (step ** group["r"]) * (lr_max ** group["weight_lr_power"])
                weight_sum = group["weight_sum"] = group["weight_sum"] + weight
                try:
                    ckp1 = weight / weight_sum
                except ZeroDivisionError:
                    ckp1 = 0
                if not group["train_mode"]:
                    raise ValueError("Not in train mode!")
            else:
                ckp1 = 0.0
            rows = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_diff"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["z"] = torch.clone(p, memory_format=torch.contiguous_format).detach()
                if "neg_pre_grad" not in st:
                    st["neg_pre_grad"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                    if step != 1:
                        # a parameter that first meets a gradient after step 1, or a resumed state without the key: the
                        # reference starts from -(clipped grad) (adan_sf.py:213-214); launch() fills it once the clip
                        # coefficient is known (the kernel only initialises it on the group's first step)
                        self._fresh.append((st["neg_pre_grad"], p.grad))
                rows.append({"p": p.detach(), "g": p.grad, "exp_avg": st["exp_avg"], "exp_avg_sq": st["exp_avg_sq"],
                             "exp_avg_diff": st["exp_avg_diff"], "z": st["z"], "neg_pre_grad": st["neg_pre_grad"],
                             "ema": next(ema_iter) if ema_iter is not None else None})
            if not rows:
                continue
            tab = self._tables.setdefault(gi, ParamTable()).build(rows)
            sf = bool(group["schedule_free"])
            if sf:
                step_size_diff = lr * (beta2 / bc2 * (1 - ckp1))
                step_size = lr * (bc1 * (1 - ckp1))
            else:
                step_size_diff = lr * beta2 / bc2
                step_size = lr / bc1
            hp = NsrAdanSF(beta1=beta1, one_minus_beta1=1 - beta1, beta2=beta2, one_minus_beta2=1 - beta2,
                           beta3=beta3, one_minus_beta3=1 - beta3, bias_correction3_sqrt=math.sqrt(bc3),
                           eps=group["eps"], decay=1 - lr * group["weight_decay"], ckp1=ckp1, step_size=step_size,
                           step_size_diff=step_size_diff, lr=lr, schedule_free=int(sf), first_step=int(step == 1),
                           max_norm=self._clip,
                           ema_lerp=float(1.0 - ema[1]) if ema is not None else 0.0,
                           ema_first=int(bool(ema[2])) if ema is not None else 0)
            hp_dev = None
            if to_device:
                # ring of pinned staging buffers: the host may run several steps ahead of the GPU, so a
                # slot is reused only after the async copy that read it has completed
                ring = self._hp_slots.get(gi)
                if ring is None:
                    n = C.sizeof(NsrAdanSF)
                    ring = self._hp_slots[gi] = {"dev": torch.zeros(n, dtype=torch.uint8, device=tab.dev.device),
                                                 "host": [torch.zeros(n, dtype=torch.uint8).pin_memory() for _ in range(4)],
                                                 "ev": [None] * 4, "i": 0}
                i = ring["i"]
                if ring["ev"][i] is not None:
                    ring["ev"][i].synchronize()
                C.memmove(ring["host"][i].data_ptr(), C.addressof(hp), C.sizeof(NsrAdanSF))
                ring["dev"].copy_(ring["host"][i], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                ring["ev"][i], ring["i"] = ev, (i + 1) % 4
                hp_dev = ring["dev"]
            self._plan.append((tab, hp, hp_dev, [r["p"] for r in rows]))

    @torch.no_grad()
    def launch(self) -> None:
        L = _lib.lib()
        sumsq_ptr = None
        if self._clip > 0:
            tab = self._all_grads_table()
            if tab.n:
                if self._sumsq is None or self._sumsq.device != tab.dev.device:
                    self._sumsq = torch.zeros(1, dtype=torch.float32, device=tab.dev.device)
                grad_sumsq(tab, self._sumsq)
                sumsq_ptr = self._sumsq.data_ptr()
        if self._fresh:  # rare, off the hot path: same coefficient the kernel derives from the squared norm
            coef = 1.0
            if sumsq_ptr is not None:
                coef = torch.clamp(self._clip / (self._sumsq.sqrt() + 1e-6), max=1.0)
            for npg, g in self._fresh:
                npg.copy_(-(g * coef))
            self._fresh = []
        for tab, hp, hp_dev, params in self._plan:
            if hp_dev is not None:
                _lib.check(L.nsr_adan_sf_step_dev(tab.dev.data_ptr(), tab.n, tab.chunks, hp_dev.data_ptr(), sumsq_ptr,
                                                  _stream()), "nsr_adan_sf_step_dev")
            else:
                _lib.check(L.nsr_adan_sf_step(tab.dev.data_ptr(), tab.n, tab.chunks, C.byref(hp), sumsq_ptr, _stream()),
                           "nsr_adan_sf_step")
            ops_mod._count(1)

    def bump_versions(self) -> None:
        for _, _, _, params in self._plan:
            torch.autograd.graph.increment_version(params)

    @torch.no_grad()
    def step(self, closure=None, *, clip_max_norm: float | None = None, ema=None):
        """ema = (list of EMA tensors aligned with this optimizer's params-with-grad, decay, first: bool)."""
        loss = closure() if closure is not None else 0.0
        self.prepare(clip_max_norm=clip_max_norm, ema=ema)
        self.launch()
        self.bump_versions()
        return loss
