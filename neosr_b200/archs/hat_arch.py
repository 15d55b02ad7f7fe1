"""HAT generator on the B200 kernels — drop-in for neosr/archs/hat_arch.py (`hat_s`, `hat_m`, `hat_l`).

Same constructor keywords, parameter / buffer names and shapes (``state_dict()`` interchanges with the
reference), `forward(x[B,3,h,w] in [0,1]) -> [B,3,s*h,s*w]`.  One explicit forward / backward over NHWC
token tensors:

  * HAB (hat_arch.py:218-350): LayerNorm -> {window self-attention (16x16 windows; roll / partition / mask
    are index math in `nsr_xwin_attn_*`)  ||  CAB: 3x3 conv, GELU, 3x3 conv, channel-attention gate} ->
    `shortcut + attn + 0.01 * conv` -> LayerNorm -> MLP;
  * OCAB (393-515): queries from 16x16 windows, keys/values from the overlapping 24x24 neighbourhood — read
    in place with zero padding, the `nn.Unfold` copy (2·C·576 floats per window) is never materialised;
  * RHAG conv, upsampler and image-side convs: the same implicit-GEMM family as SwinIR.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor, nn
from torch.nn.init import trunc_normal_

from .. import ops
from ..engine import ParamSet
from ..registry import ARCH_REGISTRY
from .arch_util import net_opt


def _rpi_sa(ws: int) -> Tensor:  # hat_arch.py:1015-1033
    ar = torch.arange(ws)
    cy, cx = torch.meshgrid(ar, ar, indexing="ij")
    cy, cx = cy.reshape(-1), cx.reshape(-1)
    return (cy[:, None] - cy[None, :] + ws - 1) * (2 * ws - 1) + (cx[:, None] - cx[None, :] + ws - 1)


def _rpi_oca(ws: int, overlap_ratio: float) -> Tensor:  # hat_arch.py:1035-1068 (entries may be negative, as there)
    wse = ws + int(overlap_ratio * ws)
    a, b = torch.arange(ws), torch.arange(wse)
    oy, ox = (t.reshape(-1) for t in torch.meshgrid(a, a, indexing="ij"))
    ey, ex = (t.reshape(-1) for t in torch.meshgrid(b, b, indexing="ij"))
    off = ws - wse + 1
    return (ey[None, :] - oy[:, None] + off) * (ws + wse - 1) + (ex[None, :] - ox[:, None] + off)


class _ChannelAttention(nn.Module):
    def __init__(self, num_feat, squeeze_factor):
        super().__init__()
        self.attention = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(num_feat, num_feat // squeeze_factor, 1, padding=0),
                                       nn.ReLU(inplace=True), nn.Conv2d(num_feat // squeeze_factor, num_feat, 1, padding=0),
                                       nn.Sigmoid())


class _CAB(nn.Module):
    def __init__(self, num_feat, compress_ratio, squeeze_factor):
        super().__init__()
        self.cab = nn.Sequential(nn.Conv2d(num_feat, num_feat // compress_ratio, 3, 1, 1), nn.GELU(),
                                 nn.Conv2d(num_feat // compress_ratio, num_feat, 3, 1, 1),
                                 _ChannelAttention(num_feat, squeeze_factor))


class _WinAttn(nn.Module):
    def __init__(self, dim, ws, heads, qkv_bias):
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) ** 2, heads))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        trunc_normal_(self.relative_position_bias_table, std=0.02)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _HAB(nn.Module):
    def __init__(self, dim, res, heads, ws, shift, compress_ratio, squeeze_factor, conv_scale, mlp_ratio, qkv_bias, drop_path):
        super().__init__()
        if min(res) <= ws:
            raise ValueError("img_size must exceed window_size (the reference then disables windows; unsupported)")
        self.shift_size, self.window_size, self.conv_scale, self.drop_prob = shift, ws, conv_scale, float(drop_path)
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _WinAttn(dim, ws, heads, qkv_bias)
        self.conv_block = _CAB(dim, compress_ratio, squeeze_factor)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _OCAB(nn.Module):
    def __init__(self, dim, ws, overlap_ratio, heads, qkv_bias, mlp_ratio):
        super().__init__()
        self.window_size = ws
        self.overlap_win_size = int(ws * overlap_ratio) + ws
        self.norm1 = nn.LayerNorm(dim)
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.relative_position_bias_table = nn.Parameter(torch.zeros((ws + self.overlap_win_size - 1) ** 2, heads))
        trunc_normal_(self.relative_position_bias_table, std=0.02)
        self.proj = nn.Linear(dim, dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _AttenBlocks(nn.Module):
    def __init__(self, blocks, ocab):
        super().__init__()
        self.blocks = nn.ModuleList(blocks)
        self.overlap_attn = ocab


class _RHAG(nn.Module):
    def __init__(self, blocks, ocab, dim):
        super().__init__()
        self.residual_group = _AttenBlocks(blocks, ocab)
        self.conv = nn.Conv2d(dim, dim, 3, 1, 1)


class _Norm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)


class hat(nn.Module):
    """Constructor mirrors neosr/archs/hat_arch.py:861-888."""

    def __init__(self, img_size=64, patch_size=1, in_chans=3, embed_dim=96, depths=(6, 6, 6, 6), num_heads=(6, 6, 6, 6),
                 window_size=7, compress_ratio=3, squeeze_factor=30, conv_scale=0.01, overlap_ratio=0.5, mlp_ratio=4.0,
                 qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1, norm_layer=nn.LayerNorm,
                 ape=False, patch_norm=True, upscale=None, img_range=1.0, upsampler="", resi_connection="1conv", **kwargs):
        super().__init__()
        if upscale is None:
            upscale = net_opt()[0]
        if patch_size != 1 or ape or drop_rate or attn_drop_rate or norm_layer is not nn.LayerNorm:
            raise NotImplementedError("neosr_b200.hat: patch_size != 1 / ape / dropout are not part of the B200 hot path")
        if upsampler != "pixelshuffle" or resi_connection != "1conv":
            raise NotImplementedError("neosr_b200.hat: upsampler='pixelshuffle' + resi_connection='1conv' (hat_s/m/l) only")
        nf = 64
        self.window_size, self.shift_size, self.overlap_ratio = window_size, window_size // 2, overlap_ratio
        self.img_range, self.upscale, self.upsampler = img_range, upscale, upsampler
        self.embed_dim, self.num_heads, self.depths = embed_dim, tuple(num_heads), tuple(depths)
        self.in_chans, self.patch_norm, self.qk_scale, self.mlp_ratio = in_chans, patch_norm, qk_scale, mlp_ratio
        self.mean = torch.full((1, 3, 1, 1), 0.5) if in_chans == 3 else torch.zeros(1, 1, 1, 1)
        self.register_buffer("relative_position_index_SA", _rpi_sa(window_size))
        self.register_buffer("relative_position_index_OCA", _rpi_oca(window_size, overlap_ratio))
        res = (img_size, img_size)
        self.conv_first = nn.Conv2d(in_chans, embed_dim, 3, 1, 1)
        self.patch_embed = _Norm(embed_dim) if patch_norm else nn.Module()
        self.layers = nn.ModuleList()
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]  # stochastic depth decay rule
        k0 = 0
        for li, depth in enumerate(depths):
            blocks = [_HAB(embed_dim, res, num_heads[li], window_size, 0 if i % 2 == 0 else window_size // 2, compress_ratio,
                           squeeze_factor, conv_scale, mlp_ratio, qkv_bias, dpr[k0 + i]) for i in range(depth)]
            k0 += depth
            ocab = _OCAB(embed_dim, window_size, overlap_ratio, num_heads[li], qkv_bias, mlp_ratio)
            self.layers.append(_RHAG(blocks, ocab, embed_dim))
        self.norm = nn.LayerNorm(embed_dim)
        self.conv_after_body = nn.Conv2d(embed_dim, embed_dim, 3, 1, 1)
        self.conv_before_upsample = nn.Sequential(nn.Conv2d(embed_dim, nf, 3, 1, 1), nn.LeakyReLU(inplace=True))
        ups = []
        if (upscale & (upscale - 1)) == 0:
            for _ in range(int(math.log2(upscale))):
                ups += [nn.Conv2d(nf, 4 * nf, 3, 1, 1), nn.PixelShuffle(2)]
        elif upscale == 3:
            ups += [nn.Conv2d(nf, 9 * nf, 3, 1, 1), nn.PixelShuffle(3)]
        else:
            raise ValueError(f"scale {upscale} is not supported. Supported scales: 2^n and 3.")
        self.upsample = nn.Sequential(*ups)
        self.conv_last = nn.Conv2d(nf, in_chans, 3, 1, 1)
        self.apply(self._init_weights)
        self._ps: ParamSet | None = None
        self._affine: dict = {}

    @staticmethod
    def _init_weights(m):  # hat_arch.py:1006-1013
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def param_set(self) -> ParamSet:
        if self._ps is None or any(self._ps._params[n] is not p for n, p in self.named_parameters()):
            self._ps = ParamSet(self)
        return self._ps

    def _consts(self, device):
        c = self._affine.get(device)
        if c is None:
            m = self.mean.to(device).flatten().expand(self.in_chans).contiguous()
            r = float(self.img_range)
            c = {"in_scale": torch.full((self.in_chans,), r, device=device), "in_shift": (-m * r).contiguous(),
                 "out_scale": torch.full((self.in_chans,), 1.0 / r, device=device), "out_shift": m.clone()}
            self._affine[device] = c
        return c

    def _drop_scale(self, blk, batch: int, device):
        """Per-sample DropPath factors for the two residual branches of a HAB (arch_util.py:118-131)."""
        if blk.drop_prob == 0.0 or not self.training:
            return None
        keep = 1.0 - blk.drop_prob
        return tuple(torch.empty(batch, device=device).bernoulli_(keep).div_(keep) for _ in range(2))

    # ------------------------------------------------------------------ explicit forward
    def engine_forward(self, x: Tensor, save: bool):
        if not x.is_cuda:
            raise RuntimeError("neosr_b200.hat runs on CUDA (sm_100a) only; there is no CPU path")
        x = x.contiguous().float()
        B, _, H, W = x.shape
        ws = self.window_size
        ows = ws + int(self.overlap_ratio * ws)
        if H % ws or W % ws:
            raise ValueError(f"input {H}x{W} must be a multiple of window_size {ws}")
        ps, k = self.param_set(), self._consts(x.device)
        ps.pack_all()  # one launch re-packs every weight image after an optimizer step
        S: dict = {"shape": (B, H, W), "layers": []} if save else None

        def lin(name, t, **kw):
            return ops.conv_fprop(t, ps.pw(name + ".weight"), ps.p(name + ".bias") if ps.has(name + ".bias") else None, **kw)

        def ln(name, t):
            return ops.layernorm_fwd(t, ps.p(name + ".weight"), ps.p(name + ".bias"))

        def mlp(pre, x1, rscale=None):
            l2, mu2, rs2 = ln(pre + "norm2", x1)
            a, hpre = lin(pre + "mlp.fc1", l2, act="gelu", want_pre=True, pre_is_actgrad=True)
            return lin(pre + "mlp.fc2", a, residual=x1, row_scale=rscale), (mu2, rs2, l2, hpre, a)

        xin = ops.nchw_to_nhwc_affine(x, k["in_scale"], k["in_shift"])
        f0 = lin("conv_first", xin)
        if self.patch_norm:
            t, mu, rs = ln("patch_embed.norm", f0)
            if save:
                S["pe"] = (mu, rs)
        else:
            t = f0
        if save:
            S["xin"], S["f0"] = xin, f0
        for li, layer in enumerate(self.layers):
            inp = t
            heads = self.num_heads[li]
            scale = self.qk_scale or (self.embed_dim // heads) ** -0.5
            blocks_saved = []
            for bi, blk in enumerate(layer.residual_group.blocks):
                pre = f"layers.{li}.residual_group.blocks.{bi}."
                cb = pre + "conv_block.cab."
                l1, mu1, rs1 = ln(pre + "norm1", t)
                qkv = lin(pre + "attn.qkv", l1)
                att, lse = ops.xwin_attn_fwd(qkv, ps.p(pre + "attn.relative_position_bias_table"), heads, ws, ws, blk.shift_size, scale)
                ds = self._drop_scale(blk, B, x.device)
                x1 = lin(pre + "attn.proj", att, residual=t, row_scale=ds[0] if ds else None)
                c1, c1g = lin(cb + "0", l1, act="gelu", want_pre=True, pre_is_actgrad=True)
                c2 = lin(cb + "2", c1)
                pooled = ops.channel_mean(c2, None, 1.0 / (H * W))
                w1, w2 = ps.p(cb + "3.attention.1.weight"), ps.p(cb + "3.attention.3.weight")
                hidden, gate = ops.channel_gate_fwd(pooled, w1.view(w1.shape[0], -1), ps.p(cb + "3.attention.1.bias"),
                                                    w2.view(w2.shape[0], -1), ps.p(cb + "3.attention.3.bias"))
                ops.channel_scale_add_(x1, c2, gate, blk.conv_scale)  # x = shortcut + attn_x + conv_x * conv_scale
                x2, ms = mlp(pre, x1, ds[1] if ds else None)
                if save:
                    blocks_saved.append((t, mu1, rs1, l1, qkv, att, lse, c1, c1g, c2, pooled, hidden, gate, x1, ms, ds))
                t = x2
            pre = f"layers.{li}.residual_group.overlap_attn."
            l1, mu1, rs1 = ln(pre + "norm1", t)
            qkv = lin(pre + "qkv", l1)
            att, lse = ops.xwin_attn_fwd(qkv, ps.p(pre + "relative_position_bias_table"), heads, ws, ows, 0, scale)
            x1 = lin(pre + "proj", att, residual=t)
            x2, ms = mlp(pre, x1)
            y = lin(f"layers.{li}.conv", x2, residual=inp)
            if save:
                S["layers"].append((blocks_saved, (t, mu1, rs1, l1, qkv, att, lse, x1, ms), x2))
            t = y
        xn, mun, rsn = ln("norm", t)
        body = lin("conv_after_body", xn, residual=f0)
        u0 = lin("conv_before_upsample.0", body, act="lrelu", act_slope=0.01)
        cur, ups = u0, []
        for i in range(len(self.upsample) // 2):
            r = self.upsample[2 * i + 1].upscale_factor
            c = lin(f"upsample.{2 * i}", cur)
            ups.append((cur, r))
            cur = ops.pixel_shuffle(c, r)
        out = lin("conv_last", cur)
        if save:
            S["final"], S["tail"] = (t, mun, rsn, xn, body), (u0, ups, cur)
        return ops.nhwc_to_nchw_affine(out, k["out_scale"], k["out_shift"]), S

    # ------------------------------------------------------------------ explicit backward
    def engine_backward(self, S: dict, dy: Tensor) -> None:
        ps = self.param_set()
        ps.ensure_grads(dy.device)
        k = self._consts(dy.device)
        ws = self.window_size
        ows = ws + int(self.overlap_ratio * ws)
        B, H, W = S["shape"]

        def bwd(name, x_in, g, need_dx=True, **epi):
            w = ps.p(name + ".weight")
            kh = w.shape[2] if w.dim() == 4 else 1
            ops.conv_wgrad(x_in, g, ps.g(name + ".weight"), ps.g(name + ".bias") if ps.has(name + ".bias") else None, kh, kh)
            return ops.conv_fprop(g, ps.pw(name + ".weight"), None, dgrad=True, **epi) if need_dx else None

        def ln_bwd(name, g, x_in, mu, rs, dres=None):
            return ops.layernorm_bwd(g, x_in, ps.p(name + ".weight"), mu, rs, ps.g(name + ".weight"), ps.g(name + ".bias"), dres=dres)

        def scaled(g, s):  # DropPath: the branch gradient is s[b] * g
            return g if s is None else (g.view(B, -1) * s.view(B, 1)).view_as(g).contiguous()

        def mlp_bwd(pre, g, x1, ms, rscale=None):
            mu2, rs2, l2, hpre, a = ms
            dh = bwd(pre + "mlp.fc2", a, scaled(g, rscale), actgrad="mulaux", aux=hpre)
            dl2 = bwd(pre + "mlp.fc1", l2, dh)
            return ln_bwd(pre + "norm2", dl2, x1, mu2, rs2, dres=g)

        g = ops.nchw_to_nhwc_affine(dy.contiguous().float(), k["out_scale"], None)
        u0, ups, last_in = S["tail"]
        g = bwd("conv_last", last_in, g)
        for i in reversed(range(len(ups))):
            src, r = ups[i]
            g = ops.pixel_unshuffle(g, r)
            g = bwd(f"upsample.{2 * i}", src, g, **(dict(actgrad="lrelu", actgrad_slope=0.01, aux=u0) if i == 0 else {}))
        t_last, mun, rsn, xn, body = S["final"]
        g = bwd("conv_before_upsample.0", body, g)
        df0 = g
        g = bwd("conv_after_body", xn, g)
        g = ln_bwd("norm", g, t_last, mun, rsn)
        for li in reversed(range(len(self.layers))):
            heads = self.num_heads[li]
            scale = self.qk_scale or (self.embed_dim // heads) ** -0.5
            blocks_saved, oc, x2 = S["layers"][li]
            dinp = g
            g = bwd(f"layers.{li}.conv", x2, g)
            # ---- OCAB
            pre = f"layers.{li}.residual_group.overlap_attn."
            t0, mu1, rs1, l1, qkv, att, lse, x1, ms = oc
            g1 = mlp_bwd(pre, g, x1, ms)
            datt = bwd(pre + "proj", att, g1)
            dqkv = ops.xwin_attn_bwd(qkv, ps.p(pre + "relative_position_bias_table"), att, datt, lse,
                                     ps.g(pre + "relative_position_bias_table"), heads, ws, ows, 0, scale)
            dl1 = bwd(pre + "qkv", l1, dqkv)
            g = ln_bwd(pre + "norm1", dl1, t0, mu1, rs1, dres=g1)
            # ---- HABs
            for bi in reversed(range(len(blocks_saved))):
                blk = self.layers[li].residual_group.blocks[bi]
                pre = f"layers.{li}.residual_group.blocks.{bi}."
                cb = pre + "conv_block.cab."
                t0, mu1, rs1, l1, qkv, att, lse, c1, c1g, c2, pooled, hidden, gate, x1, ms, ds = blocks_saved[bi]
                g1 = mlp_bwd(pre, g, x1, ms, ds[1] if ds else None)
                datt = bwd(pre + "attn.proj", att, scaled(g1, ds[0] if ds else None))
                dqkv = ops.xwin_attn_bwd(qkv, ps.p(pre + "attn.relative_position_bias_table"), att, datt, lse,
                                         ps.g(pre + "attn.relative_position_bias_table"), heads, ws, ws, blk.shift_size, scale)
                dl1 = bwd(pre + "attn.qkv", l1, dqkv)
                # conv branch: conv_x * conv_scale with conv_x = c2 * gate(mean(c2))
                dgate = ops.channel_mean(g1, c2, blk.conv_scale)
                w1, w2 = ps.p(cb + "3.attention.1.weight"), ps.p(cb + "3.attention.3.weight")
                dpooled = ops.channel_gate_bwd(dgate, gate, hidden, pooled, w1.view(w1.shape[0], -1), w2.view(w2.shape[0], -1),
                                               ps.g(cb + "3.attention.1.weight"), ps.g(cb + "3.attention.1.bias"),
                                               ps.g(cb + "3.attention.3.weight"), ps.g(cb + "3.attention.3.bias"))
                dc2 = ops.channel_scale_bwd(g1, gate, dpooled, blk.conv_scale)
                dc1 = bwd(cb + "2", c1, dc2, actgrad="mulaux", aux=c1g)
                dl1c = bwd(cb + "0", l1, dc1)
                dl1 = ops.axpby(dl1, 1.0, dl1c, 1.0, out=dl1)
                g = ln_bwd(pre + "norm1", dl1, t0, mu1, rs1, dres=g1)
            g = ops.axpby(g, 1.0, dinp, 1.0)
        if self.patch_norm:
            mu, rs = S["pe"]
            g = ln_bwd("patch_embed.norm", g, S["f0"], mu, rs, dres=df0)
        else:
            g = ops.axpby(g, 1.0, df0, 1.0)
        bwd("conv_first", S["xin"], g, need_dx=False)

    def train(self, mode: bool = True):
        if self._ps is not None:
            self._ps.invalidate_packed()
        return super().train(mode)

    def forward(self, x: Tensor) -> Tensor:
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if not need_grad:
            return self.engine_forward(x, save=False)[0]
        return _HATFn.apply(x, self, *self.parameters())


class _HATFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, net, *params):
        y, saved = net.engine_forward(x, save=True)
        ctx.net, ctx.saved = net, saved
        return y

    @staticmethod
    def backward(ctx, dy):
        net = ctx.net
        net.engine_backward(ctx.saved, dy)
        ctx.saved = None
        ps = net.param_set()
        return (None, None, *[ps.g(n) if p.requires_grad else None for n, p in net.named_parameters()])


_COMMON = dict(in_chans=3, window_size=16, conv_scale=0.01, overlap_ratio=0.5, img_range=1.0, mlp_ratio=2,
               upsampler="pixelshuffle", resi_connection="1conv")


@ARCH_REGISTRY.register()
def hat_s(**kwargs):
    return hat(compress_ratio=24, squeeze_factor=24, depths=[6] * 6, embed_dim=144, num_heads=[6] * 6, **_COMMON, **kwargs)


@ARCH_REGISTRY.register()
def hat_m(**kwargs):
    return hat(compress_ratio=3, squeeze_factor=30, depths=[6] * 6, embed_dim=180, num_heads=[6] * 6, **_COMMON, **kwargs)


@ARCH_REGISTRY.register()
def hat_l(**kwargs):
    return hat(compress_ratio=3, squeeze_factor=30, depths=[6] * 12, embed_dim=180, num_heads=[6] * 12, **_COMMON, **kwargs)
