"""SwinIR generator on the B200 kernels — drop-in for neosr/archs/swinir_arch.py.

Same registry names (`swinir_small`, `swinir_medium`, `swinir_large`), constructor keywords,
parameter/buffer names and shapes (``state_dict()`` interchanges with the reference key for
key), and `forward(x[B,3,h,w] in [0,1]) -> [B,3,s*h,s*w]`.  The implementation is not a port:
there is no per-op autograd graph.  One explicit forward and one explicit backward drive the
C-ABI kernels over NHWC token tensors:

  * roll / window_partition / window_reverse / attention mask (swinir_arch.py:41-78, 313-386)
    are index math inside `nsr_window_attn_{fwd,bwd}` — nothing is materialised;
  * Linear and Conv2d are one implicit-GEMM family with fused bias / GELU / LeakyReLU /
    residual epilogues (`nsr_conv_fprop`), dgrad is the same kernel on the rotated filter,
    wgrad is a deterministic split-K kernel;
  * PatchEmbed/PatchUnEmbed transposes (712-716, 757-761) vanish because everything stays NHWC.
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


def _rel_pos_index(ws: int) -> Tensor:
    """Buffer `relative_position_index` (swinir_arch.py:120-137); the kernels recompute it
    arithmetically, the buffer exists for state_dict compatibility."""
    ar = torch.arange(ws)
    cy, cx = torch.meshgrid(ar, ar, indexing="ij")
    cy, cx = cy.reshape(-1), cx.reshape(-1)
    return (cy[:, None] - cy[None, :] + ws - 1) * (2 * ws - 1) + (cx[:, None] - cx[None, :] + ws - 1)


def _shift_mask(h: int, w: int, ws: int, shift: int) -> Tensor:
    """Buffer `attn_mask` (swinir_arch.py:313-341) as a closed form of token coordinates."""
    def region(n):
        i = torch.arange(n)
        return (i >= n - ws).long() + (i >= n - shift).long()
    ids = (region(h)[:, None] * 3 + region(w)[None, :]).float()
    ids = ids.view(h // ws, ws, w // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = ids[:, None, :] - ids[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


class _Attn(nn.Module):
    def __init__(self, dim: int, ws: int, heads: int, qkv_bias: bool):
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) ** 2, heads))
        self.register_buffer("relative_position_index", _rel_pos_index(ws))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        trunc_normal_(self.relative_position_bias_table, std=0.02)


class _Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, res, heads, ws, shift, mlp_ratio, qkv_bias, drop_path):
        super().__init__()
        if min(res) <= ws:
            raise ValueError("img_size must exceed window_size (the reference then disables windows; unsupported)")
        self.shift_size, self.window_size, self.drop_prob = shift, ws, float(drop_path)
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _Attn(dim, ws, heads, qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        self.register_buffer("attn_mask", _shift_mask(res[0], res[1], ws, shift) if shift > 0 else None)


class _Group(nn.Module):
    def __init__(self, blocks):
        super().__init__()
        self.blocks = nn.ModuleList(blocks)


def _resi_conv(dim: int, kind: str) -> nn.Module:
    if kind == "1conv":
        return nn.Conv2d(dim, dim, 3, 1, 1)
    return nn.Sequential(nn.Conv2d(dim, dim // 4, 3, 1, 1), nn.LeakyReLU(0.2, True),
                         nn.Conv2d(dim // 4, dim // 4, 1, 1, 0), nn.LeakyReLU(0.2, True),
                         nn.Conv2d(dim // 4, dim, 3, 1, 1))


class _RSTB(nn.Module):
    def __init__(self, blocks, dim, resi):
        super().__init__()
        self.residual_group = _Group(blocks)
        self.conv = _resi_conv(dim, resi)


class _Norm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)


class swinir(nn.Module):
    """Constructor mirrors neosr/archs/swinir_arch.py:849-874."""

    def __init__(self, img_size=32, patch_size=1, in_chans=3, embed_dim=60, depths=(6, 6, 6, 6),
                 num_heads=(6, 6, 6, 6), flash_attn=False, window_size=8, mlp_ratio=2.0, qkv_bias=True,
                 qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1, norm_layer=nn.LayerNorm,
                 ape=False, patch_norm=True, use_checkpoint=False, upscale=None, img_range=1.0,
                 upsampler="pixelshuffle", resi_connection="1conv", **kwargs):
        super().__init__()
        if upscale is None:
            upscale = net_opt()[0]
        if patch_size != 1 or ape or flash_attn or drop_rate or attn_drop_rate or norm_layer is not nn.LayerNorm:
            raise NotImplementedError("neosr_b200.swinir: patch_size!=1 / ape / flash_attn / dropout are not "
                                      "part of the B200 hot path (reference defaults only)")
        if upsampler not in ("pixelshuffle", "pixelshuffledirect", "nearest+conv") or resi_connection not in ("1conv", "3conv"):
            raise NotImplementedError(f"neosr_b200.swinir: upsampler={upsampler!r}/resi={resi_connection!r} not built "
                                      "(pixelshuffle, pixelshuffledirect, nearest+conv with 1conv / 3conv are)")
        if upsampler == "nearest+conv" and upscale != 4:
            raise AssertionError("only support x4 now.")  # swinir_arch.py:993
        nf = 64
        self.img_range, self.upscale, self.upsampler = img_range, upscale, upsampler
        self.embed_dim, self.window_size, self.num_heads = embed_dim, window_size, tuple(num_heads)
        self.depths, self.patch_norm, self.qk_scale = tuple(depths), patch_norm, qk_scale
        self.in_chans, self.num_feat, self.mlp_ratio = in_chans, nf, mlp_ratio
        self.resi_connection = resi_connection
        self.mean = torch.full((1, 3, 1, 1), 0.5) if in_chans == 3 else torch.zeros(1, 1, 1, 1)
        res = (img_size, img_size)

        self.conv_first = nn.Conv2d(in_chans, embed_dim, 3, 1, 1)
        self.patch_embed = _Norm(embed_dim) if patch_norm else nn.Module()
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        k = 0
        for li, depth in enumerate(depths):
            blocks = [_Block(embed_dim, res, num_heads[li], window_size, 0 if i % 2 == 0 else window_size // 2,
                             mlp_ratio, qkv_bias, dpr[k + i]) for i in range(depth)]
            k += depth
            self.layers.append(_RSTB(blocks, embed_dim, resi_connection))
        self.norm = nn.LayerNorm(embed_dim)
        self.conv_after_body = _resi_conv(embed_dim, resi_connection)
        if upsampler == "pixelshuffle":
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
        elif upsampler == "nearest+conv":  # swinir_arch.py:991-1001
            self.conv_before_upsample = nn.Sequential(nn.Conv2d(embed_dim, nf, 3, 1, 1), nn.LeakyReLU(inplace=True))
            self.conv_up1 = nn.Conv2d(nf, nf, 3, 1, 1)
            self.conv_up2 = nn.Conv2d(nf, nf, 3, 1, 1)
            self.conv_hr = nn.Conv2d(nf, nf, 3, 1, 1)
            self.conv_last = nn.Conv2d(nf, in_chans, 3, 1, 1)
            self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        else:
            self.upsample = nn.Sequential(nn.Conv2d(embed_dim, upscale ** 2 * in_chans, 3, 1, 1),
                                          nn.PixelShuffle(upscale))
        self.apply(self._init_weights)
        self._ps: ParamSet | None = None
        self._affine: dict = {}

    @staticmethod
    def _init_weights(m):  # swinir_arch.py:1008-1015
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ------------------------------------------------------------------ engine plumbing
    def param_set(self) -> ParamSet:
        if self._ps is None or any(self._ps._params[n] is not p for n, p in self.named_parameters()):
            self._ps = ParamSet(self)
        return self._ps

    def _consts(self, device):
        c = self._affine.get(device)
        if c is None:
            m = self.mean.to(device).flatten().expand(self.in_chans).contiguous()
            r = float(self.img_range)
            c = {"in_scale": torch.full((self.in_chans,), r, device=device),
                 "in_shift": (-m * r).contiguous(),
                 "out_scale": torch.full((self.in_chans,), 1.0 / r, device=device),
                 "out_shift": m.clone()}
            self._affine[device] = c
        return c

    def _drop_scales(self, batch: int, device):
        """Per-sample DropPath factors drawn exactly like arch_util.drop_path (118-131)."""
        out = []
        for layer in self.layers:
            for blk in layer.residual_group.blocks:
                if blk.drop_prob == 0.0 or not self.training:
                    out.append(None)
                    continue
                keep = 1.0 - blk.drop_prob
                pair = []
                for _ in range(2):
                    t = torch.empty(batch, device=device).bernoulli_(keep)
                    if keep > 0.0:
                        t.div_(keep)
                    pair.append(t)
                out.append(tuple(pair))
        return out

    # ------------------------------------------------------------------ explicit forward
    def engine_forward(self, x: Tensor, save: bool):
        """x: [B,C,h,w] CUDA fp32.  Returns (y [B,C,s*h,s*w], saved-activations dict or None)."""
        if not x.is_cuda:
            raise RuntimeError("neosr_b200.swinir runs on CUDA (sm_100a) only; there is no CPU path")
        x = x.contiguous().float()
        B, _, H, W = x.shape
        ws = self.window_size
        if H % ws or W % ws:
            raise ValueError(f"input {H}x{W} must be a multiple of window_size {ws}")
        ps = self.param_set()
        ps.pack_all()  # every weight image whose parameter moved (i.e. all, after an optimizer step), in one launch
        k = self._consts(x.device)
        S: dict = {"shape": (B, H, W)} if save else None
        drops = self._drop_scales(B, x.device)
        sti = ops.sti_enabled()  # GEMM operands as split tile images (bulk-copied, no conversion warps)
        # attention operands in window order with heads padded to 32 channels: written that way by the qkv
        # contraction's epilogue and fetched by the attention kernels with bulk copies (NSR_WSTI=0: A/B switch)
        wsti = sti and ops.WSTI_ENABLED and all(ops.wsti_supported(self.embed_dim, h, ws) for h in self.num_heads)
        wsti_pad = wsti and ops.WSTI_ATTN_ENGINE != "mma_sync"  # tcgen05 attention writes a head-padded output image

        def lin(name, t, **kw):
            return ops.conv_fprop(t, ps.pw(name + ".weight"), ps.p(name + ".bias") if ps.has(name + ".bias") else None, **kw)

        def resi(name, t, residual):
            """RSTB.conv / conv_after_body: one 3x3 conv, or the 3conv bottleneck (swinir_arch.py:628-641, 962-973)."""
            if self.resi_connection == "1conv":
                return lin(name, t, residual=residual), None
            a = lin(name + ".0", t, act="lrelu", act_slope=0.2)
            b = lin(name + ".2", a, act="lrelu", act_slope=0.2)
            return lin(name + ".4", b, residual=residual), (a, b)

        xin = ops.nchw_to_nhwc_affine(x, k["in_scale"], k["in_shift"])
        f0 = lin("conv_first", xin)
        if self.patch_norm:
            t, mu, rs = ops.layernorm_fwd(f0, ps.p("patch_embed.norm.weight"), ps.p("patch_embed.norm.bias"))
            if save:
                S["pe"] = (mu, rs)
        else:
            t = f0
        if save:
            S["xin"], S["f0"], S["blocks"], S["layers"] = xin, f0, [], []
        bi_glob = 0
        for li, layer in enumerate(self.layers):
            inp = t
            heads = self.num_heads[li]
            scale = self.qk_scale or (self.embed_dim // heads) ** -0.5
            for bi, blk in enumerate(layer.residual_group.blocks):
                pre = f"layers.{li}.residual_group.blocks.{bi}."
                ds = drops[bi_glob]
                bi_glob += 1
                ln1, mu1, rs1 = ops.layernorm_fwd(t, ps.p(pre + "norm1.weight"), ps.p(pre + "norm1.bias"),
                                                  sti_out=sti, f32_out=not sti)
                if wsti:
                    qw = ps.pw_mapped(pre + "attn.qkv.weight", "qkv_rows", pre + "attn.qkv.bias",
                                      row_map=ops.head_pad_map(self.embed_dim, heads, 3))
                    qkv = ops.conv_fprop(ln1, qw, qw.bias_padded, sti_out=True, f32_out=False, sti_win=(ws, blk.shift_size))
                    att = ops.window_attn_fwd_wsti(qkv, ps.p(pre + "attn.relative_position_bias_table"), self.embed_dim,
                                                   heads, ws, blk.shift_size, scale, padded_out=wsti_pad)
                else:
                    qkv = lin(pre + "attn.qkv", ln1)
                    att = ops.window_attn_fwd(qkv, ps.p(pre + "attn.relative_position_bias_table"), heads, ws,
                                              blk.shift_size, scale, sti_out=sti)
                if wsti and wsti_pad:  # head-padded attention output: proj contracts over G channels (zero weight columns)
                    pwm = ps.pw_mapped(pre + "attn.proj.weight", "proj_cols", None,
                                       col_map=ops.head_pad_map(self.embed_dim, heads, 1))
                    x1 = ops.conv_fprop(att, pwm, ps.p(pre + "attn.proj.bias"), residual=t, row_scale=ds[0] if ds else None)
                else:
                    x1 = lin(pre + "attn.proj", att, residual=t, row_scale=ds[0] if ds else None)
                ln2, mu2, rs2 = ops.layernorm_fwd(x1, ps.p(pre + "norm2.weight"), ps.p(pre + "norm2.bias"),
                                                  sti_out=sti, f32_out=not sti)
                # hpre holds gelu'(fc1 pre-activation): the only thing the backward pass needs of it
                a, hpre = lin(pre + "mlp.fc1", ln2, act="gelu", want_pre=True, pre_is_actgrad=True, sti_out=sti,
                              f32_out=not sti, pre_u16=sti and ops.AGC_U16)
                x2 = lin(pre + "mlp.fc2", a, residual=x1, row_scale=ds[1] if ds else None)
                if save:
                    S["blocks"].append((t, mu1, rs1, ln1, qkv, att, x1, mu2, rs2, ln2, hpre, a, ds))
                t = x2
            y, rs_saved = resi(f"layers.{li}.conv", t, inp)
            if save:
                S["layers"].append((t, rs_saved))
            t = y
        xn, mun, rsn = ops.layernorm_fwd(t, ps.p("norm.weight"), ps.p("norm.bias"))
        body, cab_saved = resi("conv_after_body", xn, f0)
        if save:
            S["final"] = (t, mun, rsn, xn, body, cab_saved)
        if self.upsampler == "pixelshuffle":
            u0 = lin("conv_before_upsample.0", body, act="lrelu", act_slope=0.01)
            cur, ups = u0, []
            n_up = len(self.upsample) // 2
            for i in range(n_up):
                r = self.upsample[2 * i + 1].upscale_factor
                c = lin(f"upsample.{2 * i}", cur)
                nxt = ops.pixel_shuffle(c, r)
                ups.append((cur, r))
                cur = nxt
            out = lin("conv_last", cur)
            if save:
                S["tail"] = (u0, ups, cur)
        elif self.upsampler == "nearest+conv":  # swinir_arch.py:1056-1069
            u0 = lin("conv_before_upsample.0", body, act="lrelu", act_slope=0.01)
            n1 = ops.nearest_up2(u0)
            c1 = lin("conv_up1", n1, act="lrelu", act_slope=0.2)
            n2 = ops.nearest_up2(c1)
            c2 = lin("conv_up2", n2, act="lrelu", act_slope=0.2)
            c3 = lin("conv_hr", c2, act="lrelu", act_slope=0.2)
            out = lin("conv_last", c3)
            if save:
                S["tail"] = (u0, n1, c1, n2, c2, c3)
        else:
            c = lin("upsample.0", body)
            out = ops.pixel_shuffle(c, self.upscale)
        y = ops.nhwc_to_nchw_affine(out, k["out_scale"], k["out_shift"])
        return y, S

    # ------------------------------------------------------------------ explicit backward
    def engine_backward(self, S: dict, dy: Tensor) -> None:
        """Writes d(loss)/d(param) for every parameter into the ParamSet's flat gradient buffer
        (overwrite semantics).  dy: [B,C,s*h,s*w]."""
        ps = self.param_set()
        ps.ensure_grads(dy.device)
        k = self._consts(dy.device)
        ws = self.window_size

        sti = ops.sti_enabled() and bool(S["blocks"]) and isinstance(S["blocks"][0][3], ops.STI)
        defer = sti and ops.DEFER_WGRAD
        ln_def = dict(deferred=ps.deferred) if defer else {}  # LayerNorm dgamma / dbeta join the batched final reduction

        def split(v):  # (fp32, STI) pair or a single tensor -> (fp32 | None, STI | None)
            if isinstance(v, tuple):
                return v
            return (None, v) if isinstance(v, ops.STI) else (v, None)

        def bwd(name, x_in, g, need_dx=True, **epi):
            """wgrad (+ bias grad) of layer `name`, then dgrad.  x_in / g may be fp32 tensors, STIs or
            (fp32, STI) pairs; STI copies feed the tcgen05 bulk-copy kernels when both exist."""
            w = ps.p(name + ".weight")
            kh = w.shape[2] if w.dim() == 4 else 1
            has_b = ps.has(name + ".bias")
            xf, xs = split(x_in)
            gf, gs = split(g)
            use_sti = kh == 1 and xs is not None and gs is not None
            if use_sti and defer and (not has_b or (xs.ones and xs.shape[-1] % 64 != 0)):
                # split-K partials now, ONE reduction launch for all 1x1 weight gradients at the end of this pass
                ps.deferred.add(name, xs, gs, ps.g(name + ".weight").view(w.shape[0], -1), ps.g(name + ".bias") if has_b else None)
            else:
                ops.conv_wgrad(None if use_sti else xf, gf if (gf is not None and (has_b or not use_sti)) else None,
                               ps.g(name + ".weight"), ps.g(name + ".bias") if has_b else None, kh, kh,
                               x_sti=xs if use_sti else None, dy_sti=gs if use_sti else None)
            if need_dx:
                src = gs if (kh == 1 and gs is not None) else gf
                return ops.conv_fprop(src, ps.pw(name + ".weight"), None, dgrad=True, **epi)
            return None

        def scaled(g, s):  # DropPath: branch gradient = s[b] * g
            if s is None:
                return g
            gf, _ = split(g)
            B = gf.shape[0]
            gf = (gf.view(B, -1) * s.view(B, 1)).view_as(gf).contiguous()
            return (gf, ops.STI.from_f32(gf)) if sti else gf

        def resi_bwd(name, x_in, g, saved, **epi):
            if self.resi_connection == "1conv":
                return bwd(name, x_in, g, **epi)
            a, b = saved
            g = bwd(name + ".4", b, g, actgrad="lrelu", actgrad_slope=0.2, aux=b)
            g = bwd(name + ".2", a, g, actgrad="lrelu", actgrad_slope=0.2, aux=a)
            g = bwd(name + ".0", x_in, g)  # dim/4 input channels: may run on the narrow / exact engine (fp32 output only)
            return (g, ops.STI.from_f32(g)) if epi.get("sti_out") else g

        g = ops.nchw_to_nhwc_affine(dy.contiguous().float(), k["out_scale"], None)
        if self.upsampler == "nearest+conv":
            u0, n1, c1, n2, c2, c3 = S["tail"]
            g = bwd("conv_last", c3, g, actgrad="lrelu", actgrad_slope=0.2, aux=c3)
            g = bwd("conv_hr", c2, g, actgrad="lrelu", actgrad_slope=0.2, aux=c2)
            g = ops.nearest_up2_bwd(bwd("conv_up2", n2, g))
            g = ops.actgrad_mul(g, c1, "lrelu", 0.2)
            g = ops.nearest_up2_bwd(bwd("conv_up1", n1, g))
            g = ops.actgrad_mul(g, u0, "lrelu", 0.01)
            t_last, mun, rsn, xn, body, cab_saved = S["final"]
            g = bwd("conv_before_upsample.0", body, g)
        elif self.upsampler == "pixelshuffle":
            u0, ups, last_in = S["tail"]
            g = bwd("conv_last", last_in, g)
            for i in reversed(range(len(ups))):
                src, r = ups[i]
                g = ops.pixel_unshuffle(g, r)
                if i == 0:
                    g = bwd(f"upsample.{2 * i}", src, g, actgrad="lrelu", actgrad_slope=0.01, aux=u0)
                else:
                    g = bwd(f"upsample.{2 * i}", src, g)
            t_last, mun, rsn, xn, body, cab_saved = S["final"]
            g = bwd("conv_before_upsample.0", body, g)
        else:
            t_last, mun, rsn, xn, body, cab_saved = S["final"]
            g = ops.pixel_unshuffle(g, self.upscale)
            g = bwd("upsample.0", body, g)
        df0 = g  # through the `+ x` skip of conv_after_body (swinir_arch.py:1047)
        g = resi_bwd("conv_after_body", xn, g, cab_saved)
        g = ops.layernorm_bwd(g, t_last, ps.p("norm.weight"), mun, rsn, ps.g("norm.weight"), ps.g("norm.bias"), key="norm", **ln_def)
        nblk = len(S["blocks"])
        bi_glob = nblk
        for li in reversed(range(len(self.layers))):
            heads = self.num_heads[li]
            scale = self.qk_scale or (self.embed_dim // heads) ** -0.5
            dinp = g
            # grad w.r.t. the last block's output: fp32 for the LayerNorm backward, STI for fc2's dgrad/wgrad
            g = resi_bwd(f"layers.{li}.conv", S["layers"][li][0], g, S["layers"][li][1], sti_out=sti)
            depth = len(self.layers[li].residual_group.blocks)
            for bi in reversed(range(depth)):
                bi_glob -= 1
                pre = f"layers.{li}.residual_group.blocks.{bi}."
                t0, mu1, rs1, ln1, qkv, att, x1, mu2, rs2, ln2, hpre, a, ds = S["blocks"][bi_glob]
                shift = self.layers[li].residual_group.blocks[bi].shift_size
                # inside a residual group the token-stream gradient lives as a tile image only (ops.LN_STI_RES): the
                # LayerNorm backward of the block above reads its residual term from the image and writes no fp32 copy
                gf = split(g)[0] if split(g)[0] is not None else split(g)[1]
                gb = scaled(g, ds[1] if ds else None)
                dh = bwd(pre + "mlp.fc2", a, gb, actgrad="mulaux", aux=hpre, sti_out=sti, f32_out=not sti)
                dln2 = bwd(pre + "mlp.fc1", ln2, dh)
                lean = sti and ops.LN_STI_RES and ds is None
                g1 = ops.layernorm_bwd(dln2, x1, ps.p(pre + "norm2.weight"), mu2, rs2, ps.g(pre + "norm2.weight"),
                                       ps.g(pre + "norm2.bias"), dres=gf, sti_out=sti, f32_out=not lean, key=pre + "norm2", **ln_def)
                g1f = split(g1)[0] if split(g1)[0] is not None else split(g1)[1]
                gb = scaled(g1, ds[0] if ds else None)
                if isinstance(qkv, ops.STI):  # window-ordered operands (see engine_forward)
                    pwm = ps.pw_mapped(pre + "attn.proj.weight", "proj_cols", None,
                                       col_map=ops.head_pad_map(self.embed_dim, heads, 1))
                    if att.shape[-1] != self.embed_dim and defer:  # head-padded attention output
                        ps.deferred.add(pre + "attn.proj", att, split(gb)[1], ps.g(pre + "attn.proj.weight"),
                                        ps.g(pre + "attn.proj.bias"), col_map=pwm.col_map, bias_col=att.ones_col)
                    elif att.shape[-1] != self.embed_dim:
                        ops.conv_wgrad_mapped(att, split(gb)[1], ps.g(pre + "attn.proj.weight"), ps.g(pre + "attn.proj.bias"),
                                              pwm.col_map, att.ones_col)
                    else:
                        bwd(pre + "attn.proj", att, gb, need_dx=False)
                    datt = ops.conv_fprop(split(gb)[1], pwm, None, dgrad=True, sti_out=True, f32_out=False, sti_win=(ws, shift))
                    pad = att.shape[-1] != self.embed_dim  # tcgen05 attention kernels: head-padded images throughout
                    dqkv = ops.window_attn_bwd_wsti(qkv, ps.p(pre + "attn.relative_position_bias_table"), datt,
                                                    ps.g(pre + "attn.relative_position_bias_table"), self.embed_dim, heads,
                                                    ws, shift, scale, padded_out=pad, key=pre + "attn.bias_table", **ln_def)
                    if pad:  # qkv wgrad / dgrad on the head-padded dqkv image
                        qw = ps.pw_mapped(pre + "attn.qkv.weight", "qkv_rows", pre + "attn.qkv.bias",
                                          row_map=ops.head_pad_map(self.embed_dim, heads, 3))
                        qb = ps.g(pre + "attn.qkv.bias") if ps.has(pre + "attn.qkv.bias") else None
                        if defer:
                            ps.deferred.add(pre + "attn.qkv", split(ln1)[1], dqkv, ps.g(pre + "attn.qkv.weight"), qb,
                                            row_map=qw.row_map)
                        else:
                            ops.conv_wgrad_mapped_rows(split(ln1)[1], dqkv, ps.g(pre + "attn.qkv.weight"), qb, qw.row_map)
                        dln1 = ops.conv_fprop(dqkv, qw, None, dgrad=True)
                else:
                    datt = bwd(pre + "attn.proj", att, gb)
                    dqkv = ops.window_attn_bwd(qkv, ps.p(pre + "attn.relative_position_bias_table"), datt,
                                               ps.g(pre + "attn.relative_position_bias_table"), heads, ws, shift, scale,
                                               sti_out=sti)
                if not (isinstance(qkv, ops.STI) and att.shape[-1] != self.embed_dim):
                    dln1 = bwd(pre + "attn.qkv", ln1, dqkv)
                g = ops.layernorm_bwd(dln1, t0, ps.p(pre + "norm1.weight"), mu1, rs1, ps.g(pre + "norm1.weight"),
                                      ps.g(pre + "norm1.bias"), dres=g1f, sti_out=sti and bi > 0,
                                      f32_out=not (lean and bi > 0 and (S["blocks"][bi_glob - 1][-1] is None)),
                                      key=pre + "norm1", **ln_def)
            g = ops.axpby(split(g)[0], 1.0, dinp, 1.0)
        if self.patch_norm:
            mu, rs = S["pe"]
            g = ops.layernorm_bwd(g, S["f0"], ps.p("patch_embed.norm.weight"), mu, rs,
                                  ps.g("patch_embed.norm.weight"), ps.g("patch_embed.norm.bias"), dres=df0,
                                  key="patch_embed.norm", **ln_def)
        else:
            g = ops.axpby(g, 1.0, df0, 1.0)
        bwd("conv_first", S["xin"], g, need_dx=False)
        ps.deferred.finalize()

    # ------------------------------------------------------------------ nn.Module surface
    def train(self, mode: bool = True):
        if self._ps is not None:
            self._ps.invalidate_packed()
        return super().train(mode)

    def forward(self, x: Tensor) -> Tensor:
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if not need_grad:
            return self.engine_forward(x, save=False)[0]
        return _SwinIRFn.apply(x, self, *self.parameters())


class _SwinIRFn(torch.autograd.Function):
    """Makes the explicit engine a single autograd node so the module also works under the
    reference's own `closure` (loss.backward(), image.py:531)."""

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
        grads = [ps.g(n) if p.requires_grad else None for n, p in net.named_parameters()]
        return (None, None, *grads)


@ARCH_REGISTRY.register()
def swinir_small(**kwargs):
    return swinir(img_size=64, depths=[6, 6, 6, 6], embed_dim=60, num_heads=[6, 6, 6, 6],
                  upsampler="pixelshuffledirect", resi_connection="1conv", **kwargs)


@ARCH_REGISTRY.register()
def swinir_medium(**kwargs):
    return swinir(img_size=48, depths=[6, 6, 6, 6, 6, 6], embed_dim=180, num_heads=[6, 6, 6, 6, 6, 6],
                  upsampler="pixelshuffle", resi_connection="1conv", **kwargs)


@ARCH_REGISTRY.register()
def swinir_large(**kwargs):
    return swinir(img_size=64, embed_dim=240, depths=[6] * 9, num_heads=[8] * 9, upsampler="nearest+conv",
                  resi_connection="3conv", **kwargs)
