"""Loss factory with the reference's contract (neosr/losses/__init__.py:25-39)."""
from __future__ import annotations

from copy import deepcopy

from ..registry import LOSS_REGISTRY
from . import basic_loss, consistency_loss, gan_loss, ssim_loss, vgg_perceptual_loss  # noqa: F401


def build_loss(opt: dict):
    opt = deepcopy(opt)
    loss_type = opt.pop("type")
    return LOSS_REGISTRY.get(loss_type)(**opt)
