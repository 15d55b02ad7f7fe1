"""Pixel losses on the fused value+gradient kernels — drop-in for neosr/losses/basic_loss.py.
Each module is one autograd node: the kernel computes the loss value and d(loss)/d(pred) in a
single pass; `backward` only scales the stored gradient."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import ops
from ..registry import LOSS_REGISTRY


class _ValueGradFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, fn):
        val, grad = fn(pred.contiguous().float(), pred.requires_grad)
        ctx.grad = grad
        return val.view(())

    @staticmethod
    def backward(ctx, gout):
        g = ctx.grad
        ctx.grad = None
        return (g * gout if g is not None else None), None


@LOSS_REGISTRY.register()
class L1Loss(nn.Module):
    """basic_loss.py:24-53."""

    def __init__(self, loss_weight: float = 1.0, reduction: str = "mean") -> None:
        super().__init__()
        if reduction != "mean":
            raise NotImplementedError("neosr_b200.L1Loss: reduction='mean' only (the training configuration)")
        self.loss_weight, self.reduction = loss_weight, reduction

    def value_and_grad(self, pred: Tensor, target: Tensor, want_grad: bool = True, loss_accum: Tensor | None = None):
        return ops.l1_loss(pred, target.contiguous().float(), self.loss_weight, loss_accum, want_grad)

    def forward(self, pred: Tensor, target: Tensor, **kwargs) -> Tensor:
        return _ValueGradFn.apply(pred, lambda p, wg: self.value_and_grad(p, target, wg))


@LOSS_REGISTRY.register()
class chc_loss(nn.Module):
    """basic_loss.py:132-219 for criterion='huber', loss_lambda=0 (every reference call site)."""

    def __init__(self, loss_weight: float = 1.0, reduction: str = "mean", criterion: str = "huber",
                 loss_lambda: float = 0, clip_min: float = 0.003921, clip_max: float = 0.996078) -> None:
        super().__init__()
        if reduction != "mean" or criterion != "huber" or loss_lambda != 0:
            raise NotImplementedError("neosr_b200.chc_loss: huber criterion with loss_lambda=0 only")
        self.loss_weight, self.clip_min, self.clip_max = loss_weight, clip_min, clip_max

    def value_and_grad(self, pred, target, want_grad=True, loss_accum=None, in_scale: float = 1.0, weight=None):
        w = self.loss_weight if weight is None else weight
        return ops.charbonnier_loss(pred, target.contiguous().float(), w, loss_accum, in_scale, self.clip_min,
                                    self.clip_max, want_grad)

    def forward(self, pred: Tensor, target: Tensor, **kwargs) -> Tensor:
        return _ValueGradFn.apply(pred, lambda p, wg: self.value_and_grad(p, target, wg))
