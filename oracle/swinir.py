"""Oracle: functional restatement of the reference SwinIR generator.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function names the
reference lines it restates (paths relative to /root/reference/).  Parameters
are a plain ``dict[str, Tensor]`` keyed exactly like the reference module's
``state_dict()`` so that checkpoints interchange.

The restatement deliberately keeps the reference's *data movement* (roll,
window_partition, window_reverse, NCHW<->token transposes) so that it is an
independent check of the index math the CUDA path folds away.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F
from torch import Tensor


@dataclass
class SwinIRConfig:
    """Constructor arguments of `swinir` (neosr/archs/swinir_arch.py:849-874)."""

    img_size: int = 32
    in_chans: int = 3
    embed_dim: int = 60
    depths: tuple = (6, 6, 6, 6)
    num_heads: tuple = (6, 6, 6, 6)
    window_size: int = 8
    mlp_ratio: float = 2.0
    upscale: int = 4
    img_range: float = 1.0
    upsampler: str = "pixelshuffle"
    resi_connection: str = "1conv"
    num_feat: int = 64
    patch_norm: bool = True
    qkv_bias: bool = True
    extra: dict = field(default_factory=dict)


def swinir_medium_config(upscale: int = 4) -> SwinIRConfig:
    """neosr/archs/swinir_arch.py:1106-1116."""
    return SwinIRConfig(img_size=48, embed_dim=180, depths=(6,) * 6, num_heads=(6,) * 6,
                        upsampler="pixelshuffle", resi_connection="1conv", upscale=upscale)


def swinir_small_config(upscale: int = 4) -> SwinIRConfig:
    """neosr/archs/swinir_arch.py:1093-1103."""
    return SwinIRConfig(img_size=64, embed_dim=60, depths=(6,) * 4, num_heads=(6,) * 4,
                        upsampler="pixelshuffledirect", resi_connection="1conv", upscale=upscale)


# --------------------------------------------------------------------------- index math
def relative_position_index(ws: int) -> Tensor:
    """swinir_arch.py:120-137 — [ws*ws, ws*ws] int64 index into the (2ws-1)^2 table."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def window_partition(x: Tensor, ws: int) -> Tensor:
    """swinir_arch.py:41-57."""
    b, h, w, c = x.shape
    x = x.view(b, h // ws, ws, w // ws, ws, c)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, c)


def window_reverse(windows: Tensor, ws: int, h: int, w: int) -> Tensor:
    """swinir_arch.py:60-78."""
    b = int(windows.shape[0] / (h * w / ws / ws))
    x = windows.view(b, h // ws, w // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(b, h, w, -1)


def calculate_mask(h: int, w: int, ws: int, shift: int) -> Tensor:
    """swinir_arch.py:313-341, including its behaviour for shift == 0 (slice(-0, None)
    selects everything, so the mask comes out all-zero)."""
    img_mask = torch.zeros((1, h, w, 1))
    hs = (slice(0, -ws), slice(-ws, -shift), slice(-shift, None))
    cnt = 0
    for a in hs:
        for b in hs:
            img_mask[:, a, b, :] = cnt
            cnt += 1
    mw = window_partition(img_mask, ws).view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)


# --------------------------------------------------------------------------- blocks
def window_attention(p: dict, pre: str, x: Tensor, mask: Tensor | None, heads: int, ws: int) -> Tensor:
    """WindowAttention.forward, swinir_arch.py:150-212 (non-flash branch)."""
    b_, n, c = x.shape
    qkv = F.linear(x, p[pre + "qkv.weight"], p.get(pre + "qkv.bias"))
    qkv = qkv.reshape(b_, n, 3, heads, c // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (c // heads) ** -0.5
    attn = q @ k.transpose(-2, -1)
    idx = relative_position_index(ws).to(x.device)
    bias = p[pre + "relative_position_bias_table"][idx.view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        attn = attn.view(b_ // nw, nw, heads, n, n) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, n, n)
    attn = torch.softmax(attn, dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(b_, n, c)
    return F.linear(x, p[pre + "proj.weight"], p[pre + "proj.bias"])


def swin_block(p: dict, pre: str, x: Tensor, x_size, heads: int, ws: int, shift: int,
               input_resolution, drop_scale: tuple | None = None) -> Tensor:
    """SwinTransformerBlock.forward, swinir_arch.py:343-392.

    ``drop_scale`` = (s1, s2): optional per-sample DropPath factors [B,1,1]
    (arch_util.py:118-131) drawn by the caller; None == drop_path 0.
    """
    h, w = x_size
    b, _, c = x.shape
    if min(input_resolution) <= ws:  # swinir_arch.py:276-279
        shift, ws = 0, min(input_resolution)
    shortcut = x
    x = F.layer_norm(x, (c,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-5).view(b, h, w, c)
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    xw = window_partition(x, ws).view(-1, ws * ws, c)
    if tuple(input_resolution) == tuple(x_size):
        mask = calculate_mask(h, w, ws, shift).to(x.device) if shift > 0 else None  # registered buffer
    else:
        mask = calculate_mask(h, w, ws, shift).to(x.device)  # swinir_arch.py:371-373
    aw = window_attention(p, pre + "attn.", xw, mask, heads, ws).view(-1, ws, ws, c)
    x = window_reverse(aw, ws, h, w)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    x = x.view(b, h * w, c)
    if drop_scale is not None:
        x = x * drop_scale[0]
    x = shortcut + x
    y = F.layer_norm(x, (c,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-5)
    y = F.linear(y, p[pre + "mlp.fc1.weight"], p[pre + "mlp.fc1.bias"])  # Mlp, swinir_arch.py:32-38
    y = F.gelu(y)
    y = F.linear(y, p[pre + "mlp.fc2.weight"], p[pre + "mlp.fc2.bias"])
    if drop_scale is not None:
        y = y * drop_scale[1]
    return x + y


def _resi_conv(p: dict, pre: str, x: Tensor, kind: str) -> Tensor:
    """RSTB.conv / conv_after_body, swinir_arch.py:629-639, 963-973."""
    if kind == "1conv":
        return F.conv2d(x, p[pre + "weight"], p[pre + "bias"], 1, 1)
    x = F.leaky_relu(F.conv2d(x, p[pre + "0.weight"], p[pre + "0.bias"], 1, 1), 0.2)
    x = F.leaky_relu(F.conv2d(x, p[pre + "2.weight"], p[pre + "2.bias"], 1, 0), 0.2)
    return F.conv2d(x, p[pre + "4.weight"], p[pre + "4.bias"], 1, 1)


def forward_features(p: dict, cfg: SwinIRConfig, x: Tensor, drop_scales=None) -> Tensor:
    """swinir.forward_features, swinir_arch.py:1025-1038 with RSTB.forward 657-663."""
    b, c, h, w = x.shape
    x_size = (h, w)
    res = (cfg.img_size, cfg.img_size)
    x = x.flatten(2).transpose(1, 2)  # PatchEmbed.forward 712-716
    if cfg.patch_norm:
        x = F.layer_norm(x, (c,), p["patch_embed.norm.weight"], p["patch_embed.norm.bias"], 1e-5)
    blk_id = 0
    for li, depth in enumerate(cfg.depths):
        inp = x
        for bi in range(depth):
            shift = 0 if bi % 2 == 0 else cfg.window_size // 2
            ds = None if drop_scales is None else drop_scales[blk_id]
            x = swin_block(p, f"layers.{li}.residual_group.blocks.{bi}.", x, x_size,
                           cfg.num_heads[li], cfg.window_size, shift, res, ds)
            blk_id += 1
        y = x.transpose(1, 2).view(b, c, h, w)  # PatchUnEmbed.forward 757-761
        y = _resi_conv(p, f"layers.{li}.conv.", y, cfg.resi_connection)
        x = y.flatten(2).transpose(1, 2) + inp
    x = F.layer_norm(x, (c,), p["norm.weight"], p["norm.bias"], 1e-5)
    return x.transpose(1, 2).view(b, c, h, w)


def swinir_forward(p: dict, cfg: SwinIRConfig, x: Tensor, drop_scales=None) -> Tensor:
    """swinir.forward, swinir_arch.py:1040-1079."""
    mean = torch.full((1, 3, 1, 1), 0.5, dtype=x.dtype, device=x.device) if cfg.in_chans == 3 \
        else torch.zeros(1, 1, 1, 1, dtype=x.dtype, device=x.device)  # 880-884
    x = (x - mean) * cfg.img_range
    if cfg.upsampler == "pixelshuffle":
        x = F.conv2d(x, p["conv_first.weight"], p["conv_first.bias"], 1, 1)
        x = _resi_conv(p, "conv_after_body.", forward_features(p, cfg, x, drop_scales), cfg.resi_connection) + x
        x = F.leaky_relu(F.conv2d(x, p["conv_before_upsample.0.weight"], p["conv_before_upsample.0.bias"], 1, 1), 0.01)
        if (cfg.upscale & (cfg.upscale - 1)) == 0:  # Upsample 768-791
            for i in range(int(math.log2(cfg.upscale))):
                x = F.pixel_shuffle(F.conv2d(x, p[f"upsample.{2 * i}.weight"], p[f"upsample.{2 * i}.bias"], 1, 1), 2)
        elif cfg.upscale == 3:
            x = F.pixel_shuffle(F.conv2d(x, p["upsample.0.weight"], p["upsample.0.bias"], 1, 1), 3)
        else:
            raise ValueError(f"scale {cfg.upscale} is not supported")
        x = F.conv2d(x, p["conv_last.weight"], p["conv_last.bias"], 1, 1)
    elif cfg.upsampler == "pixelshuffledirect":
        x = F.conv2d(x, p["conv_first.weight"], p["conv_first.bias"], 1, 1)
        x = _resi_conv(p, "conv_after_body.", forward_features(p, cfg, x, drop_scales), cfg.resi_connection) + x
        x = F.pixel_shuffle(F.conv2d(x, p["upsample.0.weight"], p["upsample.0.bias"], 1, 1), cfg.upscale)
    elif cfg.upsampler == "nearest+conv":
        x = F.conv2d(x, p["conv_first.weight"], p["conv_first.bias"], 1, 1)
        x = _resi_conv(p, "conv_after_body.", forward_features(p, cfg, x, drop_scales), cfg.resi_connection) + x
        x = F.leaky_relu(F.conv2d(x, p["conv_before_upsample.0.weight"], p["conv_before_upsample.0.bias"], 1, 1), 0.01)
        x = F.leaky_relu(F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"),
                                  p["conv_up1.weight"], p["conv_up1.bias"], 1, 1), 0.2)
        x = F.leaky_relu(F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"),
                                  p["conv_up2.weight"], p["conv_up2.bias"], 1, 1), 0.2)
        x = F.leaky_relu(F.conv2d(x, p["conv_hr.weight"], p["conv_hr.bias"], 1, 1), 0.2)
        x = F.conv2d(x, p["conv_last.weight"], p["conv_last.bias"], 1, 1)
    else:
        xf = F.conv2d(x, p["conv_first.weight"], p["conv_first.bias"], 1, 1)
        r = _resi_conv(p, "conv_after_body.", forward_features(p, cfg, xf, drop_scales), cfg.resi_connection) + xf
        x = x + F.conv2d(r, p["conv_last.weight"], p["conv_last.bias"], 1, 1)
    return x / cfg.img_range + mean


# --------------------------------------------------------------------------- parameters
def swinir_param_shapes(cfg: SwinIRConfig) -> dict:
    """Names/shapes of the reference ``state_dict()`` parameters (not buffers), in
    registration order (swinir_arch.py:889-1004)."""
    c, nf, ws = cfg.embed_dim, cfg.num_feat, cfg.window_size
    hid = int(c * cfg.mlp_ratio)
    s: dict = {}

    def conv(name, co, ci, k):
        s[name + ".weight"] = (co, ci, k, k)
        s[name + ".bias"] = (co,)

    def resi(name):
        if cfg.resi_connection == "1conv":
            conv(name, c, c, 3)
        else:
            conv(name + ".0", c // 4, c, 3)
            conv(name + ".2", c // 4, c // 4, 1)
            conv(name + ".4", c, c // 4, 3)

    conv("conv_first", c, cfg.in_chans, 3)
    if cfg.patch_norm:
        s["patch_embed.norm.weight"] = (c,)
        s["patch_embed.norm.bias"] = (c,)
    for li, depth in enumerate(cfg.depths):
        for bi in range(depth):
            pre = f"layers.{li}.residual_group.blocks.{bi}."
            s[pre + "norm1.weight"] = (c,)
            s[pre + "norm1.bias"] = (c,)
            s[pre + "attn.relative_position_bias_table"] = ((2 * ws - 1) ** 2, cfg.num_heads[li])
            s[pre + "attn.qkv.weight"] = (3 * c, c)
            if cfg.qkv_bias:
                s[pre + "attn.qkv.bias"] = (3 * c,)
            s[pre + "attn.proj.weight"] = (c, c)
            s[pre + "attn.proj.bias"] = (c,)
            s[pre + "norm2.weight"] = (c,)
            s[pre + "norm2.bias"] = (c,)
            s[pre + "mlp.fc1.weight"] = (hid, c)
            s[pre + "mlp.fc1.bias"] = (hid,)
            s[pre + "mlp.fc2.weight"] = (c, hid)
            s[pre + "mlp.fc2.bias"] = (c,)
        resi(f"layers.{li}.conv")
    s["norm.weight"] = (c,)
    s["norm.bias"] = (c,)
    resi("conv_after_body")
    if cfg.upsampler == "pixelshuffle":
        conv("conv_before_upsample.0", nf, c, 3)
        if (cfg.upscale & (cfg.upscale - 1)) == 0:
            for i in range(int(math.log2(cfg.upscale))):
                conv(f"upsample.{2 * i}", 4 * nf, nf, 3)
        else:
            conv("upsample.0", 9 * nf, nf, 3)
        conv("conv_last", cfg.in_chans, nf, 3)
    elif cfg.upsampler == "pixelshuffledirect":
        conv("upsample.0", cfg.upscale ** 2 * cfg.in_chans, c, 3)
    elif cfg.upsampler == "nearest+conv":
        conv("conv_before_upsample.0", nf, c, 3)
        for n in ("conv_up1", "conv_up2", "conv_hr"):
            conv(n, nf, nf, 3)
        conv("conv_last", cfg.in_chans, nf, 3)
    else:
        conv("conv_last", cfg.in_chans, c, 3)
    return s


def synth_params(shapes: dict, seed: int = 0, dtype=torch.float32) -> dict:
    """Deterministic, *non-degenerate* synthetic parameters for parity runs: every
    tensor (including biases, LayerNorm affine and the relative-position table, which
    the reference initialises to 0/1/~0) gets seeded noise so that each gradient path
    is exercised.  Generated on CPU with torch.Generator -> identical on every box."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for k, shp in shapes.items():
        if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k.endswith("norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".bias"):
            t = 0.02 * torch.randn(shp, generator=g)
        elif k.endswith("relative_position_bias_table"):
            t = 0.2 * torch.randn(shp, generator=g)
        elif len(shp) == 4:
            fan_in = shp[1] * shp[2] * shp[3]
            t = torch.randn(shp, generator=g) * (1.0 / math.sqrt(fan_in))
        elif len(shp) == 2:
            t = torch.randn(shp, generator=g) * (1.0 / math.sqrt(shp[1]))
        else:
            t = 0.25 + 0.05 * torch.randn(shp, generator=g)  # PReLU slopes etc.
        out[k] = t.to(dtype)
    return out
