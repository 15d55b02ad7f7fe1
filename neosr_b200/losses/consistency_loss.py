"""Oklab-chroma + CIE-L* consistency loss on the fused kernels — drop-in for neosr/losses/consistency_loss.py."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import ops
from ..registry import LOSS_REGISTRY
from .basic_loss import _ValueGradFn


def _gaussian_blur_kernel(ksize: int = 21, sigma: float = 3.0) -> Tensor:
    """torchvision.transforms.GaussianBlur(21, 3)'s kernel (consistency_loss.py:45-46; torchvision
    _get_gaussian_kernel2d: linspace grid, pdf normalised in 1-D, outer product)."""
    half = (ksize - 1) * 0.5
    x = torch.linspace(-half, half, steps=ksize)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    k1 = pdf / pdf.sum()
    return torch.mm(k1[:, None], k1[None, :])


@LOSS_REGISTRY.register()
class consistency_loss(nn.Module):
    def __init__(self, criterion: str = "chc", blur: bool = True, cosim: bool = True, saturation: float = 1.0,
                 brightness: float = 1.0, loss_weight: float = 1.0) -> None:
        super().__init__()
        if criterion != "chc":
            if criterion == "l1":
                raise NotImplementedError("neosr_b200.consistency_loss: criterion 'chc' only (the template default)")
            raise NotImplementedError(f"{criterion} criterion has not been supported.")
        self.use_blur, self.cosim = blur, cosim
        self.saturation, self.brightness, self.loss_weight = saturation, brightness, loss_weight
        self.criterion_type = criterion
        self.register_buffer("_blur_kernel", _gaussian_blur_kernel(21, 3.0), persistent=False)

    def value_and_grad(self, net_output: Tensor, gt: Tensor, want_grad: bool = True, loss_accum: Tensor | None = None):
        k = self._blur_kernel if self.use_blur else None
        return ops.consistency_loss(net_output, gt.contiguous().float(), k, self.saturation, self.brightness, self.cosim,
                                    self.loss_weight, loss_accum, want_grad)

    def forward(self, net_output: Tensor, gt: Tensor) -> Tensor:
        return _ValueGradFn.apply(net_output, lambda p, wg: self.value_and_grad(p, gt, wg))
