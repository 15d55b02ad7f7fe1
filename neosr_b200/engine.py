"""Shared host-side machinery for the explicit forward/backward engines.

`ParamSet` gives an nn.Module's parameters (a) packed fprop/dgrad weight images for the
contraction kernels, re-packed only when the parameter's version changes, and (b) gradient
storage as views into ONE flat fp32 buffer, which is what the data-parallel all-reduce and the
fused optimizer consume (one all-reduce per network per step)."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from . import ops


class ParamSet:
    def __init__(self, module: nn.Module, trainable: bool = True):
        self.module = module
        self.trainable = trainable
        self.names = [n for n, _ in module.named_parameters()]
        self._params = dict(module.named_parameters())
        self._packed: dict = {}
        self._pack_tables: dict = {}
        self.deferred = ops.DeferredWgrads()  # 1x1 weight gradients reduced in one launch at the end of a backward pass
        self.flat_grad: Tensor | None = None
        self._grads: dict = {}
        self.device = None

    # -- parameters ------------------------------------------------------------------
    def p(self, name: str) -> Tensor:
        return self._params[name].detach()  # shares storage AND version counter with the Parameter

    def has(self, name: str) -> bool:
        return name in self._params

    def pw(self, name: str, need_dgrad: bool = True) -> ops.PackedWeight:
        w = self.p(name)
        pk = self._packed.get(name)
        if pk is None or pk.weight.data_ptr() != w.data_ptr():
            pk = self._packed[name] = ops.PackedWeight(w, need_dgrad=need_dgrad)
        return pk.refresh()

    def pw_mapped(self, name: str, key: str, bias_name: str | None, row_map=None, col_map=None,
                  need_dgrad: bool = True) -> ops.MappedPackedWeight:
        """A re-indexed (head-padded) packed copy of Linear weight `name`; `key` distinguishes several views."""
        w = self.p(name)
        pk = self._packed.get((name, key))
        if pk is None or pk.weight.data_ptr() != w.data_ptr():
            pk = self._packed[(name, key)] = ops.MappedPackedWeight(
                w, self.p(bias_name) if bias_name is not None and self.has(bias_name) else None, row_map, col_map, need_dgrad)
        return pk.refresh()

    def pack_all(self) -> None:
        """Re-pack EVERY packed weight whose parameter changed in ONE launch (`nsr_pack_weights_multi`).  The engines call
        this at the top of a training forward: after an optimizer step all versions have moved, and packing lazily at
        first use cost one launch per parameter (+ gathers for the head-padded copies) - ~400 five-microsecond launches per
        SwinIR-medium step.  Weights first met later (first iteration) still pack lazily in pw() / pw_mapped()."""
        stale = [pk for pk in self._packed.values() if pk.stale()]
        if len(stale) < 2:
            return  # nothing to batch; the lazy path handles it
        ops.pack_weights_multi(stale, self._pack_tables)

    def invalidate_packed(self) -> None:
        """Force a re-pack on next use (weights changed behind autograd's back, e.g. the
        schedule-free optimizer's `p.data.lerp_` in train()/eval(), adan_sf.py:112-136)."""
        for pk in self._packed.values():
            pk._version = None

    # -- gradients -------------------------------------------------------------------
    def ensure_grads(self, device) -> None:
        if self.flat_grad is not None and self.flat_grad.device == device:
            return
        total = sum(self._params[n].numel() for n in self.names)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=device)
        off = 0
        for n in self.names:
            t = self._params[n]
            self._grads[n] = self.flat_grad[off:off + t.numel()].view(t.shape)
            off += t.numel()

    def g(self, name: str) -> Tensor:
        return self._grads[name]

    def grads_in_order(self) -> list:
        return [self._grads[n] for n in self.names]

    def attach_grads(self) -> None:
        """Point every Parameter's .grad at its view of the flat buffer (what the reference's
        optimizers, clip_grad_norm_ and DDP read)."""
        for n in self.names:
            self._params[n].grad = self._grads[n]
