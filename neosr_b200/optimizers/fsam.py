"""F-SAM (Friendly Sharpness-Aware Minimization) on the fused multi-tensor kernels - drop-in for
neosr/optimizers/fsam.py: same constructor (`fsam(params, base_optimizer, rho, sigma, lmbda, adaptive, **kwargs)`),
per-parameter state names (`momentum`, `old_p`), `first_step` / `second_step` / `step(closure, current_iter)`.

The reference walks the parameters in Python (clone, sub, mul-add, pow, norm per tensor: ~10 launches per parameter and
a host-side stack of norms); here first_step is three launches over the whole network (gradient correction + momentum +
partial norms, norm, climb) and second_step is one restore launch plus the base optimizer's fused step."""
from __future__ import annotations

import torch
from torch.optim.optimizer import Optimizer

from .. import _lib
from .. import ops as ops_mod
from ..ops import _stream, scratch
from ._table import ParamTable


class fsam(Optimizer):
    def __init__(self, params, base_optimizer, rho: float = 0.5, sigma: float = 1.0, lmbda: float = 0.9,
                 adaptive: bool = True, **kwargs) -> None:
        assert rho >= 0.0, f"Invalid rho, should be non-negative: {rho}"
        defaults = dict(rho=rho, adaptive=adaptive, **kwargs)
        super().__init__(params, defaults)
        self.base_optimizer = base_optimizer(self.param_groups, **kwargs)
        self.param_groups = self.base_optimizer.param_groups
        self.defaults.update(self.base_optimizer.defaults)
        self.sigma, self.lmbda = sigma, lmbda
        self._table = ParamTable()
        self._sumsq = None

    def _build(self):
        rows = []
        first = False
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if "momentum" not in st:
                    st["momentum"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                    first = True
                if "old_p" not in st:
                    st["old_p"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                rows.append({"p": p.detach(), "g": p.grad, "exp_avg": st["momentum"], "z": st["old_p"]})
        return self._table.build(rows), first, rows

    @torch.no_grad()
    def first_step(self, zero_grad: bool = False) -> None:  # fsam.py:36-68
        tab, first, rows = self._build()
        if not tab.n:
            return
        g0 = self.param_groups[0]
        L = _lib.lib()
        dev = tab.dev.device
        if self._sumsq is None or self._sumsq.device != dev:
            self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        ws = scratch(L.nsr_grad_sumsq_workspace(), dev)
        _lib.check(L.nsr_fsam_first_step(tab.dev.data_ptr(), tab.n, tab.chunks, float(g0["rho"]), float(self.sigma),
                                         float(self.lmbda), int(bool(g0["adaptive"])), int(first), self._sumsq.data_ptr(),
                                         ws.data_ptr(), _stream()), "nsr_fsam_first_step")
        ops_mod._count(3)
        torch.autograd.graph.increment_version([r["p"] for r in rows])  # the weights moved: packed images are stale
        if zero_grad:
            self.zero_grad(set_to_none=True)

    @torch.no_grad()
    def second_step(self, zero_grad: bool = False, **step_kw) -> None:  # fsam.py:70-80
        tab, _, rows = self._build()
        if tab.n:
            _lib.check(_lib.lib().nsr_fsam_restore(tab.dev.data_ptr(), tab.n, tab.chunks, _stream()), "nsr_fsam_restore")
            ops_mod._count(1)
            torch.autograd.graph.increment_version([r["p"] for r in rows])
        self.base_optimizer.step(**step_kw)  # the actual "sharpness-aware" update
        if zero_grad:
            self.zero_grad(set_to_none=True)

    @torch.no_grad()
    def step(self, closure=None, current_iter: int | None = None):
        assert closure is not None, "Sharpness Aware Minimization requires closure, but it was not provided"
        closure = torch.enable_grad()(closure)
        self.first_step(zero_grad=True)
        closure(current_iter)
        self.second_step()

    def load_state_dict(self, state_dict: dict) -> None:
        super().load_state_dict(state_dict)
        self.base_optimizer.param_groups = self.param_groups
