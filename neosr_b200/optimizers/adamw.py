"""torch.optim.AdamW semantics on the fused multi-tensor kernel (base.py:154-155 selects AdamW
for `optim_g.type = "adamw"`, config C5).  State names (`step`, `exp_avg`, `exp_avg_sq`) follow
torch so optimizer checkpoints interchange."""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch.optim.optimizer import Optimizer

from .. import _lib
from .._lib import NsrAdamW
from .. import ops as ops_mod
from ..ops import _stream
from ._table import ParamTable, grad_sumsq


class AdamW(Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 amsgrad: bool = False, **kwargs) -> None:
        if amsgrad:
            raise NotImplementedError("neosr_b200.AdamW: amsgrad not built")
        super().__init__(params, {"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay,
                                  "amsgrad": False})
        self._tables: dict = {}
        self._sumsq = None
        self._hp_slots: dict = {}
        self._plan: list = []
        self._clip = 0.0

    # step() = prepare() (host: advance `step`, derive the bias corrections) + launch() (device: grad-norm + ONE fused
    # kernel) + bump_versions(); in CUDA-graph mode the model calls prepare(to_device=True) every iteration (one small
    # async copy per group) and replays a graph that contains launch() - the same split as adan_sf.
    @torch.no_grad()
    def prepare(self, *, clip_max_norm: float | None = None, ema=None, to_device: bool = False) -> None:
        self._plan = []
        self._clip = float(clip_max_norm or 0.0)
        ema_iter = iter(ema[0]) if ema is not None else None
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            rows, step = [], None
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                if step is None:
                    step = int(st["step"].item())  # host tensor: no device sync
                rows.append({"p": p.detach(), "g": p.grad, "exp_avg": st["exp_avg"], "exp_avg_sq": st["exp_avg_sq"],
                             "ema": next(ema_iter) if ema_iter is not None else None})
            if not rows:
                continue
            tab = self._tables.setdefault(gi, ParamTable()).build(rows)
            bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
            hp = NsrAdamW(beta1=b1, one_minus_beta1=1 - b1, beta2=b2, one_minus_beta2=1 - b2, eps=group["eps"],
                          decay=1 - group["lr"] * group["weight_decay"], step_size=group["lr"] / bc1,
                          bias_correction2_sqrt=math.sqrt(bc2), max_norm=self._clip,
                          ema_lerp=float(1.0 - ema[1]) if ema is not None else 0.0,
                          ema_first=int(bool(ema[2])) if ema is not None else 0)
            hp_dev = None
            if to_device:  # ring of pinned staging buffers, as in adan_sf.prepare
                ring = self._hp_slots.get(gi)
                if ring is None:
                    n = C.sizeof(NsrAdamW)
                    ring = self._hp_slots[gi] = {"dev": torch.zeros(n, dtype=torch.uint8, device=tab.dev.device),
                                                 "host": [torch.zeros(n, dtype=torch.uint8).pin_memory() for _ in range(4)],
                                                 "ev": [None] * 4, "i": 0}
                i = ring["i"]
                if ring["ev"][i] is not None:
                    ring["ev"][i].synchronize()
                C.memmove(ring["host"][i].data_ptr(), C.addressof(hp), C.sizeof(NsrAdamW))
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
            rows = [{"p": p.detach(), "g": p.grad} for g in self.param_groups for p in g["params"] if p.grad is not None]
            tab = self._tables.setdefault("norm", ParamTable()).build(rows)
            if self._sumsq is None:
                self._sumsq = torch.zeros(1, dtype=torch.float32, device=tab.dev.device)
            grad_sumsq(tab, self._sumsq)
            sumsq_ptr = self._sumsq.data_ptr()
        for tab, hp, hp_dev, _ in self._plan:
            if hp_dev is not None:
                _lib.check(L.nsr_adamw_step_dev(tab.dev.data_ptr(), tab.n, tab.chunks, hp_dev.data_ptr(), sumsq_ptr, _stream()),
                           "nsr_adamw_step_dev")
            else:
                _lib.check(L.nsr_adamw_step(tab.dev.data_ptr(), tab.n, tab.chunks, C.byref(hp), sumsq_ptr, _stream()),
                           "nsr_adamw_step")
            ops_mod._count(1)

    def bump_versions(self) -> None:
        for _, _, _, params in self._plan:
            torch.autograd.graph.increment_version(params)

    @torch.no_grad()
    def step(self, closure=None, *, clip_max_norm: float | None = None, ema=None):
        loss = closure() if closure is not None else None
        self.prepare(clip_max_norm=clip_max_norm, ema=ema)
        self.launch()
        self.bump_versions()
        return loss
