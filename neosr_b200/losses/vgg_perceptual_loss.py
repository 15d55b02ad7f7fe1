"""VGG19 perceptual loss on the B200 kernels — drop-in for
neosr/losses/vgg_perceptual_loss.py:57-242 (patchloss/ipk branch excluded).

forward(x, gt): features of x and gt, per-tap `criterion(fx/10, fg/10) * layer_weight`, summed,
times loss_weight (204-242).  Value and d(loss)/dx are produced in ONE forward sweep:
VGG fprop(x) saving ReLU outputs, VGG fprop(gt), fused charbonnier value+grad per tap, VGG
dgrad chain back to the image."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import ops
from ..archs.vgg_arch import VGGFeatureExtractor
from ..registry import LOSS_REGISTRY


class _PerceptualFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gt, mod):
        val, dx = mod.value_and_grad(x, gt, want_grad=x.requires_grad)
        ctx.dx = dx
        return val.view(())

    @staticmethod
    def backward(ctx, gout):
        dx = ctx.dx
        ctx.dx = None
        return (dx * gout if dx is not None else None), None, None


@LOSS_REGISTRY.register()
class vgg_perceptual_loss(nn.Module):
    def __init__(self, layer_weights=None, vgg_type: str = "vgg19", use_input_norm: bool = True,
                 range_norm: bool = False, loss_weight: float = 1.0, criterion: str = "chc", patchloss: bool = False,
                 ipk: bool = False, patch_weight: float = 1.0, allow_random_init: bool | None = None, **kwargs) -> None:
        super().__init__()
        if patchloss or ipk:
            raise NotImplementedError("neosr_b200.vgg_perceptual_loss: PatchLoss/IPK branch not built")
        if criterion != "chc":
            raise NotImplementedError("neosr_b200.vgg_perceptual_loss: criterion 'chc' (the template default) only")
        self.loss_weight = loss_weight
        self.layer_weights = dict(layer_weights) if layer_weights is not None else {
            "conv1_2": 0.1, "conv2_2": 0.1, "conv3_4": 1.0, "conv4_4": 1.0, "conv5_4": 1.0}
        self.vgg = VGGFeatureExtractor(layer_name_list=list(self.layer_weights.keys()), vgg_type=vgg_type,
                                       use_input_norm=use_input_norm, range_norm=range_norm,
                                       allow_random_init=allow_random_init)
        self.criterion_type = criterion

    def value_and_grad(self, x: Tensor, gt: Tensor, want_grad: bool = True, loss_accum: Tensor | None = None):
        """Returns (loss value [1] device tensor, d(loss)/dx [B,3,H,W] or None)."""
        fx, S = self.vgg.engine_forward(x, save=want_grad)
        fg, _ = self.vgg.engine_forward(gt.detach(), save=False)
        total = torch.zeros(1, dtype=torch.float32, device=x.device)
        dtaps = {}
        for k, w in self.layer_weights.items():
            # chc_loss(loss_lambda=0, clip 0..1) on fx/10, fg/10 (vgg_perceptual_loss.py:145,232-236)
            _, d = ops.charbonnier_loss(fx[k], fg[k], float(w) * self.loss_weight, total, in_scale=0.1,
                                        clip_min=0.0, clip_max=1.0, want_grad=want_grad)
            dtaps[k] = d
        if loss_accum is not None:
            ops.axpby(total, 1.0, loss_accum, 1.0, out=loss_accum)
        dx = self.vgg.engine_backward(S, dtaps) if want_grad else None
        return total, dx

    def forward(self, x: Tensor, gt: Tensor) -> Tensor:
        return _PerceptualFn.apply(x, gt, self)
