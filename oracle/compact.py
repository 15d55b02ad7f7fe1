"""Oracle: SRVGGNetCompact (`compact`) forward, neosr/archs/compact_arch.py:76-85 (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import torch.nn.functional as F
from torch import Tensor


def compact_param_shapes(num_in_ch=3, num_out_ch=3, num_feat=64, num_conv=16, upscale=4, act_type="prelu") -> dict:
    """state_dict parameter names/shapes in registration order (compact_arch.py:47-73): `body` is a
    ModuleList alternating conv / activation, so conv k sits at body[2k] and its PReLU at body[2k+1]."""
    s = {}
    cin = num_in_ch
    for k in range(num_conv + 1):
        s[f"body.{2 * k}.weight"] = (num_feat, cin, 3, 3)
        s[f"body.{2 * k}.bias"] = (num_feat,)
        if act_type == "prelu":
            s[f"body.{2 * k + 1}.weight"] = (num_feat,)
        cin = num_feat
    last = 2 * (num_conv + 1)
    s[f"body.{last}.weight"] = (num_out_ch * upscale * upscale, num_feat, 3, 3)
    s[f"body.{last}.bias"] = (num_out_ch * upscale * upscale,)
    return s


def compact_forward(p: dict, x: Tensor, num_conv=16, upscale=4, act_type="prelu") -> Tensor:
    out = x
    for k in range(num_conv + 1):
        out = F.conv2d(out, p[f"body.{2 * k}.weight"], p[f"body.{2 * k}.bias"], 1, 1)
        if act_type == "prelu":
            out = F.prelu(out, p[f"body.{2 * k + 1}.weight"])
        elif act_type == "relu":
            out = F.relu(out)
        else:
            out = F.leaky_relu(out, 0.1)
    last = 2 * (num_conv + 1)
    out = F.conv2d(out, p[f"body.{last}.weight"], p[f"body.{last}.bias"], 1, 1)
    out = F.pixel_shuffle(out, upscale)
    return out + F.interpolate(x, scale_factor=upscale, mode="nearest")
