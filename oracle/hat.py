"""Oracle: functional restatement of the reference HAT generator (neosr/archs/hat_arch.py).  TEST INFRASTRUCTURE.

Parameters are a ``dict[str, Tensor]`` keyed like the reference ``state_dict()``.  Keeps the reference's data
movement (roll, window_partition/reverse, nn.Unfold for the overlapping keys/values, indexing the bias table with
the registered — partly negative — relative_position_index_OCA) so it checks the index math the CUDA path folds
away.  Pinned against the live module in tests/test_oracle_vs_reference.py and fixtures in tests/golden/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import Tensor

from oracle.swinir import calculate_mask, window_partition, window_reverse


@dataclass
class HATConfig:
    """hat.__init__ arguments (hat_arch.py:861-888); hat_l values by default (1190-1207)."""
    img_size: int = 64
    in_chans: int = 3
    embed_dim: int = 180
    depths: tuple = (6,) * 12
    num_heads: tuple = (6,) * 12
    window_size: int = 16
    compress_ratio: int = 3
    squeeze_factor: int = 30
    conv_scale: float = 0.01
    overlap_ratio: float = 0.5
    mlp_ratio: float = 2
    upscale: int = 4
    img_range: float = 1.0
    num_feat: int = 64


def rpi_sa(ws: int) -> Tensor:
    """hat.calculate_rpi_sa, hat_arch.py:1015-1033."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def rpi_oca(ws: int, overlap_ratio: float) -> Tensor:
    """hat.calculate_rpi_oca, hat_arch.py:1035-1068."""
    wse = ws + int(overlap_ratio * ws)
    co = torch.flatten(torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij")), 1)
    ce = torch.flatten(torch.stack(torch.meshgrid([torch.arange(wse), torch.arange(wse)], indexing="ij")), 1)
    rel = (ce[:, None, :] - co[:, :, None]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - wse + 1
    rel[:, :, 1] += ws - wse + 1
    rel[:, :, 0] *= ws + wse - 1
    return rel.sum(-1)


def _lin(p, name, x):
    return F.linear(x, p[name + ".weight"], p.get(name + ".bias"))


def _ln(p, name, x):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], 1e-5)


def _mlp(p, pre, x):
    return _lin(p, pre + "fc2", F.gelu(_lin(p, pre + "fc1", x)))


def cab(p: dict, pre: str, x: Tensor) -> Tensor:
    """CAB + ChannelAttention, hat_arch.py:15-52; x is NCHW."""
    y = F.conv2d(x, p[pre + "cab.0.weight"], p[pre + "cab.0.bias"], 1, 1)
    y = F.conv2d(F.gelu(y), p[pre + "cab.2.weight"], p[pre + "cab.2.bias"], 1, 1)
    a = F.adaptive_avg_pool2d(y, 1)
    a = F.relu(F.conv2d(a, p[pre + "cab.3.attention.1.weight"], p[pre + "cab.3.attention.1.bias"]))
    a = torch.sigmoid(F.conv2d(a, p[pre + "cab.3.attention.3.weight"], p[pre + "cab.3.attention.3.bias"]))
    return y * a


def window_attention(p: dict, pre: str, x: Tensor, rpi: Tensor, mask, heads: int, ws: int) -> Tensor:
    """WindowAttention.forward, hat_arch.py:168-215."""
    b_, n, c = x.shape
    qkv = _lin(p, pre + "qkv", x).reshape(b_, n, 3, heads, c // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (c // heads) ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = p[pre + "relative_position_bias_table"][rpi.view(-1)].view(ws * ws, ws * ws, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        attn = (attn.view(b_ // nw, nw, heads, n, n) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, n, n)
    attn = torch.softmax(attn, dim=-1)
    return _lin(p, pre + "proj", (attn @ v).transpose(1, 2).reshape(b_, n, c))


def hab(p: dict, pre: str, x: Tensor, x_size, rpi: Tensor, attn_mask: Tensor, heads: int, ws: int, shift: int,
        conv_scale: float, drop_scale=None) -> Tensor:
    """HAB.forward, hat_arch.py:299-350.  drop_scale: optional per-sample DropPath factors (attn, mlp)."""
    h, w = x_size
    b, _, c = x.shape
    shortcut = x
    x = _ln(p, pre + "norm1", x).view(b, h, w, c)
    conv_x = cab(p, pre + "conv_block.", x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).contiguous().view(b, h * w, c)
    sx = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2)) if shift > 0 else x
    xw = window_partition(sx, ws).view(-1, ws * ws, c)
    aw = window_attention(p, pre + "attn.", xw, rpi, attn_mask if shift > 0 else None, heads, ws).view(-1, ws, ws, c)
    sx = window_reverse(aw, ws, h, w)
    ax = (torch.roll(sx, shifts=(shift, shift), dims=(1, 2)) if shift > 0 else sx).view(b, h * w, c)
    if drop_scale is not None:
        ax = ax * drop_scale[0].view(b, 1, 1)
    x = shortcut + ax + conv_x * conv_scale
    m = _mlp(p, pre + "mlp.", _ln(p, pre + "norm2", x))
    if drop_scale is not None:
        m = m * drop_scale[1].view(b, 1, 1)
    return x + m


def ocab(p: dict, pre: str, x: Tensor, x_size, rpi: Tensor, heads: int, ws: int, ows: int) -> Tensor:
    """OCAB.forward, hat_arch.py:445-515 (einops rearrange written out)."""
    h, w = x_size
    b, _, c = x.shape
    shortcut = x
    x = _ln(p, pre + "norm1", x).view(b, h, w, c)
    qkv = _lin(p, pre + "qkv", x).reshape(b, h, w, 3, c).permute(3, 0, 4, 1, 2)  # 3, b, c, h, w
    q = qkv[0].permute(0, 2, 3, 1)
    kv = torch.cat((qkv[1], qkv[2]), dim=1)
    qw = window_partition(q, ws).view(-1, ws * ws, c)
    kvw = F.unfold(kv, kernel_size=(ows, ows), stride=ws, padding=(ows - ws) // 2)  # b, 2c*ows*ows, nw
    nw = kvw.shape[-1]
    kvw = kvw.view(b, 2, c, ows, ows, nw).permute(1, 0, 5, 3, 4, 2).reshape(2, b * nw, ows * ows, c)
    kw_, vw = kvw[0], kvw[1]
    b_, nq, _ = qw.shape
    n, d = kw_.shape[1], c // heads
    q = qw.reshape(b_, nq, heads, d).permute(0, 2, 1, 3) * d ** -0.5
    k = kw_.reshape(b_, n, heads, d).permute(0, 2, 1, 3)
    v = vw.reshape(b_, n, heads, d).permute(0, 2, 1, 3)
    attn = q @ k.transpose(-2, -1)
    bias = p[pre + "relative_position_bias_table"][rpi.view(-1)].view(ws * ws, ows * ows, -1).permute(2, 0, 1).contiguous()
    attn = torch.softmax(attn + bias.unsqueeze(0), dim=-1)
    aw = (attn @ v).transpose(1, 2).reshape(b_, nq, c).view(-1, ws, ws, c)
    x = window_reverse(aw, ws, h, w).view(b, h * w, c)
    x = _lin(p, pre + "proj", x) + shortcut
    return x + _mlp(p, pre + "mlp.", _ln(p, pre + "norm2", x))


def hat_forward(p: dict, cfg: HATConfig, x: Tensor, drop_scales=None) -> Tensor:
    """hat.forward / forward_features / RHAG / AttenBlocks, hat_arch.py:607-616,711-720,1109-1147."""
    ws, ows = cfg.window_size, cfg.window_size + int(cfg.overlap_ratio * cfg.window_size)
    mean = torch.full((1, 3, 1, 1), 0.5, dtype=x.dtype)
    x = (x - mean) * cfg.img_range
    f0 = F.conv2d(x, p["conv_first.weight"], p["conv_first.bias"], 1, 1)
    h, w = f0.shape[2:]
    mask = calculate_mask(h, w, ws, ws // 2).to(x.dtype)
    ri_sa, ri_oca = rpi_sa(ws), rpi_oca(ws, cfg.overlap_ratio)
    t = _ln(p, "patch_embed.norm", f0.flatten(2).transpose(1, 2))
    bi_glob = 0
    for li, depth in enumerate(cfg.depths):
        inp = t
        for bi in range(depth):
            ds = drop_scales[bi_glob] if drop_scales is not None else None
            bi_glob += 1
            t = hab(p, f"layers.{li}.residual_group.blocks.{bi}.", t, (h, w), ri_sa, mask, cfg.num_heads[li], ws,
                    0 if bi % 2 == 0 else ws // 2, cfg.conv_scale, ds)
        t = ocab(p, f"layers.{li}.residual_group.overlap_attn.", t, (h, w), ri_oca, cfg.num_heads[li], ws, ows)
        y = t.transpose(1, 2).view(-1, cfg.embed_dim, h, w)
        y = F.conv2d(y, p[f"layers.{li}.conv.weight"], p[f"layers.{li}.conv.bias"], 1, 1)
        t = y.flatten(2).transpose(1, 2) + inp
    t = _ln(p, "norm", t).transpose(1, 2).view(-1, cfg.embed_dim, h, w)
    y = F.conv2d(t, p["conv_after_body.weight"], p["conv_after_body.bias"], 1, 1) + f0
    y = F.leaky_relu(F.conv2d(y, p["conv_before_upsample.0.weight"], p["conv_before_upsample.0.bias"], 1, 1), 0.01)
    if (cfg.upscale & (cfg.upscale - 1)) == 0:
        for i in range(int(math.log2(cfg.upscale))):
            y = F.pixel_shuffle(F.conv2d(y, p[f"upsample.{2 * i}.weight"], p[f"upsample.{2 * i}.bias"], 1, 1), 2)
    else:
        y = F.pixel_shuffle(F.conv2d(y, p["upsample.0.weight"], p["upsample.0.bias"], 1, 1), 3)
    y = F.conv2d(y, p["conv_last.weight"], p["conv_last.bias"], 1, 1)
    return y / cfg.img_range + mean


def hat_param_shapes(cfg: HATConfig) -> dict:
    c, nf, ws = cfg.embed_dim, cfg.num_feat, cfg.window_size
    ows = ws + int(cfg.overlap_ratio * ws)
    hid = int(c * cfg.mlp_ratio)
    s: dict = {}

    def conv(name, co, ci, k):
        s[name + ".weight"], s[name + ".bias"] = (co, ci, k, k), (co,)

    def lin(name, co, ci):
        s[name + ".weight"], s[name + ".bias"] = (co, ci), (co,)

    def norm(name):
        s[name + ".weight"], s[name + ".bias"] = (c,), (c,)

    conv("conv_first", c, cfg.in_chans, 3)
    norm("patch_embed.norm")
    for li, depth in enumerate(cfg.depths):
        heads = cfg.num_heads[li]
        for bi in range(depth):
            pre = f"layers.{li}.residual_group.blocks.{bi}."
            norm(pre + "norm1")
            s[pre + "attn.relative_position_bias_table"] = ((2 * ws - 1) ** 2, heads)
            lin(pre + "attn.qkv", 3 * c, c)
            lin(pre + "attn.proj", c, c)
            conv(pre + "conv_block.cab.0", c // cfg.compress_ratio, c, 3)
            conv(pre + "conv_block.cab.2", c, c // cfg.compress_ratio, 3)
            conv(pre + "conv_block.cab.3.attention.1", c // cfg.squeeze_factor, c, 1)
            conv(pre + "conv_block.cab.3.attention.3", c, c // cfg.squeeze_factor, 1)
            norm(pre + "norm2")
            lin(pre + "mlp.fc1", hid, c)
            lin(pre + "mlp.fc2", c, hid)
        pre = f"layers.{li}.residual_group.overlap_attn."
        norm(pre + "norm1")
        lin(pre + "qkv", 3 * c, c)
        s[pre + "relative_position_bias_table"] = ((ws + ows - 1) ** 2, heads)
        lin(pre + "proj", c, c)
        norm(pre + "norm2")
        lin(pre + "mlp.fc1", hid, c)
        lin(pre + "mlp.fc2", c, hid)
        conv(f"layers.{li}.conv", c, c, 3)
    norm("norm")
    conv("conv_after_body", c, c, 3)
    conv("conv_before_upsample.0", nf, c, 3)
    if (cfg.upscale & (cfg.upscale - 1)) == 0:
        for i in range(int(math.log2(cfg.upscale))):
            conv(f"upsample.{2 * i}", 4 * nf, nf, 3)
    else:
        conv("upsample.0", 9 * nf, nf, 3)
    conv("conv_last", cfg.in_chans, nf, 3)
    return s
