"""MS-SSIM loss on the fused kernels — drop-in for neosr/losses/ssim_loss.py (`mssim_loss`, `GaussianFilter2D`)."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import ops
from ..registry import LOSS_REGISTRY
from .basic_loss import _ValueGradFn


class GaussianFilter2D(nn.Module):
    """Holds the reference's `gaussian_window` buffer ([C,1,k,k], ssim_loss.py:11-65); the filtering itself is
    fused into the SSIM kernels."""

    def __init__(self, window_size: int = 11, in_channels: int = 3, sigma: float = 1.5, padding: int | None = None) -> None:
        super().__init__()
        if window_size % 2 != 1:
            raise ValueError("Window size must be odd.")
        self.window_size, self.sigma = window_size, sigma
        self.padding = padding if padding is not None else window_size // 2
        if self.padding != window_size // 2:
            raise NotImplementedError("neosr_b200.GaussianFilter2D: padding = window_size // 2 only (the default)")
        x = torch.arange(-(window_size // 2), window_size // 2 + 1)
        w = torch.exp(-0.5 * x**2 / (sigma * sigma))
        w = (w / w.sum()).reshape(1, 1, 1, window_size)
        k2 = torch.matmul(w.transpose(-1, -2), w)
        self.register_buffer("gaussian_window", k2.repeat(in_channels, 1, 1, 1))


@LOSS_REGISTRY.register()
class mssim_loss(nn.Module):
    def __init__(self, window_size: int = 11, in_channels: int = 3, sigma: float = 1.5, K1: float = 0.01, K2: float = 0.03,
                 L: int = 1, padding: int | None = None, loss_weight: float = 1.0) -> None:
        super().__init__()
        self.window_size = window_size
        self.C1, self.C2 = (K1 * L) ** 2, (K2 * L) ** 2
        self.loss_weight = loss_weight
        self.gaussian_filter = GaussianFilter2D(window_size=window_size, in_channels=in_channels, sigma=sigma, padding=padding)

    def value_and_grad(self, x: Tensor, y: Tensor, want_grad: bool = True, loss_accum: Tensor | None = None):
        assert x.shape == y.shape, f"x: {x.shape} and y: {y.shape} must be the same"
        assert x.ndim == y.ndim == 4, f"x: {x.ndim} and y: {y.ndim} must be 4"
        win = self.gaussian_filter.gaussian_window[0, 0].contiguous()
        return ops.msssim_loss(x, y.contiguous().float(), win, self.loss_weight, self.C1, self.C2, loss_accum, want_grad)

    def forward(self, x: Tensor, y: Tensor) -> Tensor:
        return _ValueGradFn.apply(x, lambda p, wg: self.value_and_grad(p, y, wg))
