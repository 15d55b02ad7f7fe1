"""Oracle: loss stack of the reference `image.closure` (TEST INFRASTRUCTURE).

Restates neosr/losses/basic_loss.py, vgg_perceptual_loss.py, gan_loss.py and
neosr/archs/vgg_arch.py as pure functions over tensors.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor

# torchvision vgg19 "E" configuration == neosr/archs/vgg_arch.py:11-49 NAMES["vgg19"].
VGG19_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M")
VGG19_NAMES = (
    "conv1_1", "relu1_1", "conv1_2", "relu1_2", "pool1",
    "conv2_1", "relu2_1", "conv2_2", "relu2_2", "pool2",
    "conv3_1", "relu3_1", "conv3_2", "relu3_2", "conv3_3", "relu3_3", "conv3_4", "relu3_4", "pool3",
    "conv4_1", "relu4_1", "conv4_2", "relu4_2", "conv4_3", "relu4_3", "conv4_4", "relu4_4", "pool4",
    "conv5_1", "relu5_1", "conv5_2", "relu5_2", "conv5_3", "relu5_3", "conv5_4", "relu5_4", "pool5",
)
DEFAULT_LAYER_WEIGHTS = {"conv1_2": 0.1, "conv2_2": 0.1, "conv3_4": 1.0, "conv4_4": 1.0, "conv5_4": 1.0}


def vgg19_conv_shapes(max_layer: str = "conv5_4") -> dict:
    """Shapes keyed like the reference extractor's state_dict: ``vgg_net.<name>.weight``
    (vgg_arch.py:136-149 renames torchvision's features[i] to NAMES[i])."""
    out, cin = {}, 3
    stop = VGG19_NAMES.index(max_layer)
    for i, name in enumerate(VGG19_NAMES[: stop + 1]):
        if name.startswith("conv"):
            stage = int(name[4])
            cout = (64, 128, 256, 512, 512)[stage - 1]
            out[f"vgg_net.{name}.weight"] = (cout, cin, 3, 3)
            out[f"vgg_net.{name}.bias"] = (cout,)
            cin = cout
    return out


def vgg19_features(p: dict, x: Tensor, layer_names) -> dict:
    """VGGFeatureExtractor.forward, vgg_arch.py:176-199, with use_input_norm=True
    (mean .5 / std .25, vgg_arch.py:168-174) and range_norm=False; taps are taken at the
    conv outputs, i.e. *before* the ReLU (the reference lists conv*_* names)."""
    x = (x - 0.5) / 0.25
    stop = max(VGG19_NAMES.index(n) for n in layer_names)
    out = {}
    for name in VGG19_NAMES[: stop + 1]:
        if name.startswith("conv"):
            x = F.conv2d(x, p[f"vgg_net.{name}.weight"], p[f"vgg_net.{name}.bias"], 1, 1)
        elif name.startswith("relu"):
            x = F.relu(x)
        else:
            x = F.max_pool2d(x, kernel_size=2, stride=2)
        if name in layer_names:
            out[name] = x.clone()
    return out


def l1_loss(pred: Tensor, target: Tensor, loss_weight: float = 1.0) -> Tensor:
    """L1Loss.forward, basic_loss.py:45-53 (reduction='mean')."""
    return loss_weight * (pred - target).abs().mean()


def mse_loss(pred: Tensor, target: Tensor, loss_weight: float = 1.0) -> Tensor:
    return loss_weight * F.mse_loss(pred, target)


def chc_loss(pred: Tensor, target: Tensor, loss_weight: float = 1.0, criterion: str = "huber",
             loss_lambda: float = 0.0, clip_min: float = 0.003921, clip_max: float = 0.996078) -> Tensor:
    """chc_loss.forward, basic_loss.py:180-219."""
    cos = (1 - F.cosine_similarity(pred, target, dim=1, eps=1e-20)).mean()
    if criterion == "l1":
        v = (pred - target).abs() + loss_lambda * cos
    else:
        v = torch.sqrt((pred - target) ** 2 + 1e-12) + loss_lambda * cos
    return loss_weight * torch.clamp(v, clip_min, clip_max).mean()


def vgg_perceptual_loss(vgg_p: dict, x: Tensor, gt: Tensor, loss_weight: float = 1.0,
                        layer_weights: dict | None = None, criterion: str = "chc") -> Tensor:
    """vgg_perceptual_loss.forward, vgg_perceptual_loss.py:204-242 (patchloss=False)."""
    lw = layer_weights or DEFAULT_LAYER_WEIGHTS
    fx = vgg19_features(vgg_p, x, list(lw))
    fg = vgg19_features(vgg_p, gt.detach(), list(lw))
    total = 0.0
    for k in fx:
        a, b = fx[k] / 10, fg[k] / 10
        if criterion == "chc":
            c = chc_loss(a, b, loss_lambda=0, clip_min=0, clip_max=1)  # :145
        elif criterion == "l1":
            c = F.l1_loss(a, b)
        elif criterion == "l2":
            c = F.mse_loss(a, b)
        else:
            c = F.huber_loss(a, b)
        total = total + c * lw[k]
    return total * loss_weight


def gan_loss(logits: Tensor, target_is_real: bool, is_disc: bool, gan_type: str = "bce",
             loss_weight: float = 0.1, real_label_val: float = 1.0, fake_label_val: float = 0.0) -> Tensor:
    """gan_loss.forward, gan_loss.py:62-82."""
    tgt = torch.full_like(logits, real_label_val if target_is_real else fake_label_val)
    if gan_type == "bce":
        loss = F.binary_cross_entropy_with_logits(logits, tgt)
    elif gan_type == "mse":
        loss = F.mse_loss(logits, tgt)
    else:
        loss = F.huber_loss(logits, tgt)
    return loss if is_disc else loss * loss_weight
