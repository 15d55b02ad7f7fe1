"""GAN loss (BCE-with-logits vs a constant label) — drop-in for neosr/losses/gan_loss.py:6-82."""
from __future__ import annotations

from torch import Tensor, nn

from .. import ops
from ..registry import LOSS_REGISTRY
from .basic_loss import _ValueGradFn


@LOSS_REGISTRY.register()
class gan_loss(nn.Module):
    def __init__(self, gan_type: str = "bce", real_label_val: float = 1.0, fake_label_val: float = 0.0,
                 loss_weight: float = 0.1) -> None:
        super().__init__()
        if gan_type != "bce":
            raise NotImplementedError(f"neosr_b200.gan_loss: gan_type {gan_type!r} not built (bce is)")
        self.gan_type, self.loss_weight = gan_type, loss_weight
        self.real_label_val, self.fake_label_val = real_label_val, fake_label_val

    def value_and_grad(self, logits: Tensor, target_is_real: bool, is_disc: bool, want_grad=True, loss_accum=None):
        label = self.real_label_val if target_is_real else self.fake_label_val
        w = 1.0 if is_disc else self.loss_weight  # loss_weight only for generators (gan_loss.py:81-82)
        return ops.bce_logits_loss(logits, label, w, loss_accum, want_grad)

    def forward(self, net_output: Tensor, target_is_real: bool, is_disc: bool = False) -> Tensor:
        return _ValueGradFn.apply(net_output, lambda p, wg: self.value_and_grad(p, target_is_real, is_disc, wg))
