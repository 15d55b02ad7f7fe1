"""Oracle: U-Net discriminator with spectral norm, neosr/archs/unet_arch.py:40-67 plus the
torch.nn.utils.spectral_norm semantics it relies on (one power iteration per training forward,
u/v updated in place and treated as constants by autograd).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's CPU legs.  Pinned against the live reference module in
tests/test_oracle_vs_reference.py."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor

SN_CONVS = ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "conv8")


def unet_param_shapes(num_in_ch=3, num_feat=64) -> tuple[dict, dict]:
    nf = num_feat
    p = {"conv0.weight": (nf, num_in_ch, 3, 3), "conv0.bias": (nf,)}
    geo = {"conv1": (2 * nf, nf, 4), "conv2": (4 * nf, 2 * nf, 4), "conv3": (8 * nf, 4 * nf, 4), "conv4": (4 * nf, 8 * nf, 3),
           "conv5": (2 * nf, 4 * nf, 3), "conv6": (nf, 2 * nf, 3), "conv7": (nf, nf, 3), "conv8": (nf, nf, 3)}
    b = {}
    for n, (co, ci, k) in geo.items():
        p[n + ".weight_orig"] = (co, ci, k, k)
        b[n + ".weight_u"] = (co,)
        b[n + ".weight_v"] = (ci * k * k,)
    p["conv9.weight"] = (1, nf, 3, 3)
    p["conv9.bias"] = (1,)
    return p, b


def synth_unet(num_in_ch=3, num_feat=64, seed=0) -> tuple[dict, dict]:
    g = torch.Generator().manual_seed(seed)
    ps, bs = unet_param_shapes(num_in_ch, num_feat)
    p = {}
    for n, s in ps.items():
        if n.endswith("bias"):
            p[n] = torch.randn(s, generator=g) * 0.05
        else:
            fan_in = s[1] * s[2] * s[3]
            p[n] = torch.randn(s, generator=g) * (1.0 / fan_in) ** 0.5
    b = {n: F.normalize(torch.randn(s, generator=g), dim=0, eps=1e-12) for n, s in bs.items()}
    return p, b


def sn_weight(w: Tensor, u: Tensor, v: Tensor, training: bool, eps: float = 1e-12) -> Tensor:
    """torch/nn/utils/spectral_norm.py compute_weight: power iteration under no_grad, in place on u/v."""
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=eps))
    uu, vv = u.clone(), v.clone()
    sigma = torch.dot(uu, torch.mv(wm, vv))
    return w / sigma


def unet_forward(p: dict, b: dict, x: Tensor, training: bool = True, skip_connection: bool = True) -> Tensor:
    w = {n: sn_weight(p[n + ".weight_orig"], b[n + ".weight_u"], b[n + ".weight_v"], training) for n in SN_CONVS}
    lr = lambda t: F.leaky_relu(t, 0.2)  # noqa: E731
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)  # noqa: E731
    x0 = lr(F.conv2d(x, p["conv0.weight"], p["conv0.bias"], 1, 1))
    x1 = lr(F.conv2d(x0, w["conv1"], None, 2, 1))
    x2 = lr(F.conv2d(x1, w["conv2"], None, 2, 1))
    x3 = lr(F.conv2d(x2, w["conv3"], None, 2, 1))
    x4 = lr(F.conv2d(up(x3), w["conv4"], None, 1, 1))
    if skip_connection:
        x4 = x4 + x2
    x5 = lr(F.conv2d(up(x4), w["conv5"], None, 1, 1))
    if skip_connection:
        x5 = x5 + x1
    x6 = lr(F.conv2d(up(x5), w["conv6"], None, 1, 1))
    if skip_connection:
        x6 = x6 + x0
    out = lr(F.conv2d(x6, w["conv7"], None, 1, 1))
    out = lr(F.conv2d(out, w["conv8"], None, 1, 1))
    return F.conv2d(out, p["conv9.weight"], p["conv9.bias"], 1, 1)
