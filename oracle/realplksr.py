"""Oracle: RealPLKSR forward, neosr/archs/realplksr_arch.py:14-23 (DCCM), 27-41 (PLKConv2d, training branch),
44-53 (EA), 56-99 (PLKBlock), 103-162 (realplksr).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's CPU legs.  Pinned against the live reference module in
tests/test_oracle_vs_reference.py."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor


def realplksr_param_shapes(in_ch=3, out_ch=3, dim=64, n_blocks=28, upscaling_factor=4, kernel_size=17, split_ratio=0.25,
                           use_ea=True) -> dict:
    s = {"feats.0.weight": (dim, in_ch, 3, 3), "feats.0.bias": (dim,)}
    pdim = int(dim * split_ratio)
    for i in range(1, n_blocks + 1):
        p = f"feats.{i}."
        s[p + "channel_mixer.0.weight"] = (2 * dim, dim, 3, 3)
        s[p + "channel_mixer.0.bias"] = (2 * dim,)
        s[p + "channel_mixer.2.weight"] = (dim, 2 * dim, 3, 3)
        s[p + "channel_mixer.2.bias"] = (dim,)
        s[p + "lk.conv.weight"] = (pdim, pdim, kernel_size, kernel_size)
        s[p + "lk.conv.bias"] = (pdim,)
        if use_ea:
            s[p + "attn.f.0.weight"] = (dim, dim, 3, 3)
            s[p + "attn.f.0.bias"] = (dim,)
        s[p + "refine.weight"] = (dim, dim, 1, 1)
        s[p + "refine.bias"] = (dim,)
        s[p + "norm.weight"] = (dim,)
        s[p + "norm.bias"] = (dim,)
    last = f"feats.{n_blocks + 2}."  # index n_blocks+1 is the Dropout2d
    s[last + "weight"] = (out_ch * upscaling_factor ** 2, dim, 3, 3)
    s[last + "bias"] = (out_ch * upscaling_factor ** 2,)
    return s


def realplksr_forward(p: dict, x: Tensor, n_blocks=28, upscaling_factor=4, kernel_size=17, split_ratio=0.25, use_ea=True,
                      norm_groups=4) -> Tensor:
    dim = p["feats.0.weight"].shape[0]
    pdim = int(dim * split_ratio)
    f = F.conv2d(x, p["feats.0.weight"], p["feats.0.bias"], 1, 1)
    for i in range(1, n_blocks + 1):
        q = f"feats.{i}."
        skip = f
        f = F.conv2d(f, p[q + "channel_mixer.0.weight"], p[q + "channel_mixer.0.bias"], 1, 1)
        f = F.mish(f)
        f = F.conv2d(f, p[q + "channel_mixer.2.weight"], p[q + "channel_mixer.2.bias"], 1, 1)
        x1, x2 = torch.split(f, [pdim, dim - pdim], dim=1)
        x1 = F.conv2d(x1, p[q + "lk.conv.weight"], p[q + "lk.conv.bias"], 1, kernel_size // 2)
        f = torch.cat([x1, x2], dim=1)
        if use_ea:
            f = f * torch.sigmoid(F.conv2d(f, p[q + "attn.f.0.weight"], p[q + "attn.f.0.bias"], 1, 1))
        f = F.conv2d(f, p[q + "refine.weight"], p[q + "refine.bias"])
        f = F.group_norm(f, norm_groups, p[q + "norm.weight"], p[q + "norm.bias"], 1e-5)
        f = f + skip
    last = f"feats.{n_blocks + 2}."
    f = F.conv2d(f, p[last + "weight"], p[last + "bias"], 1, 1)
    f = f + torch.repeat_interleave(x, upscaling_factor ** 2, dim=1)
    return F.pixel_shuffle(f, upscaling_factor)
