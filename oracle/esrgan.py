"""Oracle: RRDBNet (`esrgan`) forward, neosr/archs/esrgan_arch.py:60-79,109-116,137-142,196-214
(TEST INFRASTRUCTURE)."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor


def esrgan_param_shapes(num_in_ch=3, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32) -> dict:
    if scale == 2:
        num_in_ch *= 4
    elif scale == 1:
        num_in_ch *= 16
    s = {}

    def conv(name, co, ci):
        s[name + ".weight"] = (co, ci, 3, 3)
        s[name + ".bias"] = (co,)

    conv("conv_first", num_feat, num_in_ch)
    for b in range(num_block):
        for r in (1, 2, 3):
            pre = f"body.{b}.rdb{r}."
            for k in range(1, 5):
                conv(pre + f"conv{k}", num_grow_ch, num_feat + (k - 1) * num_grow_ch)
            conv(pre + "conv5", num_feat, num_feat + 4 * num_grow_ch)
    for n in ("conv_body", "conv_up1", "conv_up2", "conv_hr"):
        conv(n, num_feat, num_feat)
    conv("conv_last", num_out_ch, num_feat)
    return s


def pixel_unshuffle(x: Tensor, scale: int) -> Tensor:
    """esrgan_arch.py:60-79."""
    b, c, hh, hw = x.size()
    h, w = hh // scale, hw // scale
    x = x.view(b, c, h, scale, w, scale)
    return x.permute(0, 1, 3, 5, 2, 4).reshape(b, c * scale ** 2, h, w)


def _c(p, name, x):
    return F.conv2d(x, p[name + ".weight"], p[name + ".bias"], 1, 1)


def esrgan_forward(p: dict, x: Tensor, scale=4, num_block=23) -> Tensor:
    feat = pixel_unshuffle(x, 2) if scale == 2 else (pixel_unshuffle(x, 4) if scale == 1 else x)
    feat = _c(p, "conv_first", feat)
    t = feat
    for b in range(num_block):
        rin = t
        for r in (1, 2, 3):
            pre = f"body.{b}.rdb{r}."
            x0 = t
            x1 = F.leaky_relu(_c(p, pre + "conv1", x0), 0.2)
            x2 = F.leaky_relu(_c(p, pre + "conv2", torch.cat((x0, x1), 1)), 0.2)
            x3 = F.leaky_relu(_c(p, pre + "conv3", torch.cat((x0, x1, x2), 1)), 0.2)
            x4 = F.leaky_relu(_c(p, pre + "conv4", torch.cat((x0, x1, x2, x3), 1)), 0.2)
            x5 = _c(p, pre + "conv5", torch.cat((x0, x1, x2, x3, x4), 1))
            t = x5 * 0.2 + x0
        t = t * 0.2 + rin
    feat = feat + _c(p, "conv_body", t)
    feat = F.leaky_relu(_c(p, "conv_up1", F.interpolate(feat, scale_factor=2, mode="nearest")), 0.2)
    feat = F.leaky_relu(_c(p, "conv_up2", F.interpolate(feat, scale_factor=2, mode="nearest")), 0.2)
    return _c(p, "conv_last", F.leaky_relu(_c(p, "conv_hr", feat), 0.2))
