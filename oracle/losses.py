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


# ---------------------------------------------------------------- MS-SSIM (ssim_loss.py:66-163)
def gaussian_window(window_size: int = 11, sigma: float = 1.5) -> Tensor:
    """GaussianFilter2D._get_gaussian_window1d/2d, ssim_loss.py:42-52 (one channel, [k,k])."""
    x = torch.arange(-(window_size // 2), window_size // 2 + 1)
    w = torch.exp(-0.5 * x**2 / (sigma * sigma))
    w = (w / w.sum()).reshape(1, window_size)
    return torch.matmul(w.t(), w)


def msssim_loss(x: Tensor, y: Tensor, loss_weight: float = 1.0, window_size: int = 11, sigma: float = 1.5,
                K1: float = 0.01, K2: float = 0.03, L: float = 1.0) -> Tensor:
    """mssim_loss.forward / msssim / _ssim, ssim_loss.py:112-163."""
    C1, C2 = (K1 * L) ** 2, (K2 * L) ** 2
    win = gaussian_window(window_size, sigma).to(x.dtype).view(1, 1, window_size, window_size).repeat(x.shape[1], 1, 1, 1)
    filt = lambda t: F.conv2d(t, win, stride=1, padding=window_size // 2, groups=t.shape[1])  # noqa: E731
    comps = []
    for i, w in enumerate((0.0448, 0.2856, 0.3001, 0.2363, 0.1333)):
        mu_x, mu_y = filt(x), filt(y)
        s2x, s2y, sxy = filt(x * x) - mu_x * mu_x, filt(y * y) - mu_y * mu_y, filt(x * y) - mu_x * mu_y
        A1, A2 = 2 * mu_x * mu_y + C1, 2 * sxy + C2
        B1, B2 = mu_x.pow(2) + mu_y.pow(2) + C1, s2x + s2y + C2
        cs = A2 / B2
        ssim = (A1 / B1) * cs
        if i == 4:
            comps.append(ssim.mean() ** w)
        else:
            comps.append(cs.mean() ** w)
            pad = [s % 2 for s in x.shape[2:]]
            x, y = F.avg_pool2d(x, 2, 2, padding=pad), F.avg_pool2d(y, 2, 2, padding=pad)
    prod = comps[0]
    for c in comps[1:]:
        prod = prod * c
    return loss_weight * (1 - prod)


# ---------------------------------------------------------------- consistency (consistency_loss.py:14-192)
def _lin_rgb(img: Tensor) -> Tensor:
    return torch.where(img <= 0.04045, img / 12.92, torch.pow((img + 0.055) / 1.055, 2.4))


def _oklab_chroma(img: Tensor) -> Tensor:
    img = _lin_rgb(img)
    r, g, b = img[:, 0], img[:, 1], img[:, 2]
    l = 0.4122214708 * r + 0.5363325363 * g + 0.0514459929 * b
    m = 0.2119034982 * r + 0.6806995451 * g + 0.1073969566 * b
    s = 0.0883024619 * r + 0.2817188376 * g + 0.6299787005 * b
    l_, m_, s_ = (t.sign() * t.abs().pow(1 / 3) for t in (l, m, s))
    a = 1.9779984951 * l_ - 2.4285922050 * m_ + 0.4505937099 * s_
    bb = 0.0259040371 * l_ + 0.7827717662 * m_ - 0.8086757660 * s_
    return torch.stack([a, bb], dim=1)


def _l_star(img: Tensor) -> Tensor:
    img = _lin_rgb(img.permute(0, 2, 3, 1)) @ torch.tensor([0.2126, 0.7152, 0.0722], dtype=img.dtype)
    img = torch.where(img <= (216 / 24389), img * (img * (24389 / 27)), img.sign() * img.abs().pow(1 / 3) * 116 - 16)
    return torch.clamp(img / 100, 0, 1)


def gaussian_blur_21_3(x: Tensor) -> Tensor:
    """torchvision GaussianBlur(21, 3): reflect pad 10, separable kernel as a 2-D depthwise conv."""
    half = 10.0
    t = torch.linspace(-half, half, steps=21)
    pdf = torch.exp(-0.5 * (t / 3.0).pow(2))
    k1 = pdf / pdf.sum()
    k2 = torch.mm(k1[:, None], k1[None, :]).to(x.dtype)
    c = x.shape[1]
    return F.conv2d(F.pad(x, (10, 10, 10, 10), mode="reflect"), k2.expand(c, 1, 21, 21), groups=c)


def _chc(a: Tensor, b: Tensor) -> Tensor:  # chc_loss(loss_lambda=0, clip_min=0, clip_max=1), basic_loss.py:192-219
    return torch.mean(torch.clamp(torch.sqrt((a - b) ** 2 + 1e-12), 0, 1))


def consistency_loss(x: Tensor, gt: Tensor, loss_weight: float = 1.0, blur: bool = True, cosim: bool = True,
                     saturation: float = 1.0, brightness: float = 1.0, force_cosim: bool | None = None) -> Tensor:
    """consistency_loss.forward, consistency_loss.py:146-192 (criterion chc).  force_cosim overrides the
    data-dependent `cosim < 1e-3` branch (tests only)."""
    x, gt = torch.clamp(x, 1 / 255, 1), torch.clamp(gt, 1 / 255, 1)
    if blur:
        il = _l_star(torch.clamp(gaussian_blur_21_3(x), 0, 1))
        tl = _l_star(torch.clamp(gaussian_blur_21_3(gt), 0, 1)) * brightness
    else:
        il, tl = _l_star(x), _l_star(gt) * brightness
    ic = torch.clamp(_oklab_chroma(x) + 0.5, 0, 1)
    tc = torch.clamp(_oklab_chroma(gt) * saturation + 0.5, 0, 1)
    loss = _chc(il, tl) + _chc(ic, tc)
    if cosim:
        sim = torch.nn.CosineSimilarity(dim=1, eps=1e-20)
        cs = 0.5 * (1 - sim(ic, tc).mean()) + 0.5 * (1 - sim(il, tl).mean())
        if (force_cosim is None and cs < 1e-3) or force_cosim:
            loss = loss + cs
    return loss * loss_weight
