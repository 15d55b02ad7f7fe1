"""RealPLKSR on the B200 kernels — drop-in for neosr/archs/realplksr_arch.py:103-167 (`realplksr`,
`realplksr_s`; same constructor keywords and `feats.N....` state_dict keys).

Per PLKBlock (realplksr_arch.py:56-99): DCCM = conv3x3(64->128) -> Mish -> conv3x3(128->64); the partial
large-kernel conv touches only the first `pdim` channels (a channel-slab view, no split/cat); EA gate
x * sigmoid(conv3x3(x)); 1x1 refine; GroupNorm fused with the block's skip add.  DySample and
Dropout2d(p > 0) are not built."""
from __future__ import annotations

import torch
from torch import Tensor, nn
from torch.nn.init import trunc_normal_

from .. import ops
from ..engine import ParamSet
from ..registry import ARCH_REGISTRY
from .arch_util import net_opt


class _DCCM(nn.Sequential):
    def __init__(self, dim):
        super().__init__(nn.Conv2d(dim, dim * 2, 3, 1, 1), nn.Mish(), nn.Conv2d(dim * 2, dim, 3, 1, 1))
        trunc_normal_(self[-1].weight, std=0.02)


class _PLKConv2d(nn.Module):
    def __init__(self, dim, kernel_size):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim, kernel_size, 1, kernel_size // 2)
        trunc_normal_(self.conv.weight, std=0.02)
        self.idx = dim


class _EA(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.f = nn.Sequential(nn.Conv2d(dim, dim, 3, 1, 1), nn.Sigmoid())
        trunc_normal_(self.f[0].weight, std=0.02)


class _PLKBlock(nn.Module):
    def __init__(self, dim, kernel_size, split_ratio, norm_groups, use_ea=True):
        super().__init__()
        self.channel_mixer = _DCCM(dim)
        self.lk = _PLKConv2d(int(dim * split_ratio), kernel_size)
        self.attn = _EA(dim) if use_ea else nn.Identity()
        self.refine = nn.Conv2d(dim, dim, 1, 1, 0)
        trunc_normal_(self.refine.weight, std=0.02)
        self.norm = nn.GroupNorm(norm_groups, dim)


@ARCH_REGISTRY.register()
class realplksr(nn.Module):
    def __init__(self, in_ch=3, out_ch=3, dim=64, n_blocks=28, upscaling_factor=None, kernel_size=17, split_ratio=0.25,
                 use_ea=True, norm_groups=4, dropout=0, dysample=False, **kwargs):
        super().__init__()
        if upscaling_factor is None:
            upscaling_factor = net_opt()[0]
        if dysample:
            raise NotImplementedError("neosr_b200.realplksr: dysample upsampler not built (pixelshuffle is)")
        if dropout:
            raise NotImplementedError("neosr_b200.realplksr: Dropout2d(p > 0) not built")
        self.upscale, self.dim, self.n_blocks = upscaling_factor, dim, n_blocks
        self.ks, self.pdim, self.use_ea, self.groups = kernel_size, int(dim * split_ratio), use_ea, norm_groups
        self.in_ch, self.out_ch = in_ch, out_ch
        self.feats = nn.Sequential(
            *[nn.Conv2d(in_ch, dim, 3, 1, 1)]
            + [_PLKBlock(dim, kernel_size, split_ratio, norm_groups, use_ea) for _ in range(n_blocks)]
            + [nn.Dropout2d(0)]
            + [nn.Conv2d(dim, out_ch * upscaling_factor ** 2, 3, 1, 1)])
        trunc_normal_(self.feats[0].weight, std=0.02)
        trunc_normal_(self.feats[-1].weight, std=0.02)
        self._ps: ParamSet | None = None

    def param_set(self) -> ParamSet:
        if self._ps is None or any(self._ps._params[n] is not p for n, p in self.named_parameters()):
            self._ps = ParamSet(self)
        return self._ps

    def train(self, mode: bool = True):
        if self._ps is not None:
            self._ps.invalidate_packed()
        return super().train(mode)

    # ------------------------------------------------------------------ forward
    def engine_forward(self, x: Tensor, save: bool):
        if not x.is_cuda:
            raise RuntimeError("neosr_b200.realplksr runs on CUDA (sm_100a) only; there is no CPU path")
        ps = self.param_set()
        ps.pack_all()  # one launch re-packs every weight image after an optimizer step
        dim, pdim = self.dim, self.pdim

        def conv(name, src, **kw):
            return ops.conv_fprop(src, ps.pw(name + ".weight"), ps.p(name + ".bias"), **kw)

        xin = ops.nchw_to_nhwc_affine(x.contiguous().float(), None, None)
        f = conv("feats.0", xin)
        blocks = []
        for i in range(1, self.n_blocks + 1):
            q = f"feats.{i}."
            skip = f
            h1 = conv(q + "channel_mixer.0", f)
            a1 = ops.mish_fwd(h1)
            h2 = conv(q + "channel_mixer.2", a1)
            # partial large-kernel conv: channels [0, pdim) replaced, the rest passes through (:35-41)
            t = torch.empty_like(h2)
            conv(q + "lk.conv", ops.Slab(h2, 0, pdim), out=ops.Slab(t, 0, pdim))
            ops.axpby2d(ops.Slab(h2, pdim, dim - pdim), 1.0, None, 0.0, out=ops.Slab(t, pdim, dim - pdim))
            if self.use_ea:
                s = conv(q + "attn.f.0", t)
                u = ops.mul_sigmoid_fwd(t, s)
            else:
                s, u = None, t
            r = conv(q + "refine", u)
            f, mean, rstd = ops.groupnorm_fwd(r, ps.p(q + "norm.weight"), ps.p(q + "norm.bias"), self.groups, 1e-5,
                                              residual=skip)
            if save:
                blocks.append((skip, h1, a1, h2, t, s, u, r, mean, rstd))
        last = f"feats.{self.n_blocks + 2}"
        o = conv(last, f)
        ops.add_repeat_interleave_(o, xin, self.upscale ** 2)       # feats(x) + repeat_interleave(x) (:158)
        if self.upscale > 1:
            o = ops.pixel_shuffle(o, self.upscale)
        y = ops.nhwc_to_nchw_affine(o, None, None)
        return y, ({"xin": xin, "blocks": blocks, "f": f} if save else None)

    # ------------------------------------------------------------------ backward
    def engine_backward(self, S: dict, dy: Tensor) -> None:
        ps = self.param_set()
        ps.ensure_grads(dy.device)
        dim, pdim, k = self.dim, self.pdim, self.ks

        def wgrad(name, x_in, g, kk=3):
            ops.conv_wgrad(x_in, g, ps.g(name + ".weight"), ps.g(name + ".bias"), kk, kk)

        def dgrad(name, g, **kw):
            return ops.conv_fprop(g, ps.pw(name + ".weight"), None, dgrad=True, **kw)

        g = ops.nchw_to_nhwc_affine(dy.contiguous().float(), None, None)
        if self.upscale > 1:
            g = ops.pixel_unshuffle(g, self.upscale)
        last = f"feats.{self.n_blocks + 2}"
        wgrad(last, S["f"], g)
        g = dgrad(last, g)
        for i in range(self.n_blocks, 0, -1):
            q = f"feats.{i}."
            skip, h1, a1, h2, t, s, u, r, mean, rstd = S["blocks"][i - 1]
            dr = ops.groupnorm_bwd(g, r, ps.p(q + "norm.weight"), mean, rstd, ps.g(q + "norm.weight"), ps.g(q + "norm.bias"),
                                   self.groups)
            wgrad(q + "refine", u, dr, 1)
            du = dgrad(q + "refine", dr)
            if self.use_ea:
                dt, ds = ops.mul_sigmoid_bwd(du, t, s)
                wgrad(q + "attn.f.0", t, ds)
                dt = dgrad(q + "attn.f.0", ds, residual=dt)
            else:
                dt = du
            dt1 = ops.axpby2d(ops.Slab(dt, 0, pdim), 1.0, None, 0.0)      # dense copy of the first pdim channels
            ops.conv_wgrad(ops.Slab(h2, 0, pdim), dt1, ps.g(q + "lk.conv.weight"), ps.g(q + "lk.conv.bias"), k, k)
            ops.conv_fprop(dt1, ps.pw(q + "lk.conv.weight"), None, dgrad=True, out=ops.Slab(dt, 0, pdim))  # dt becomes dh2
            wgrad(q + "channel_mixer.2", a1, dt)
            dh1 = ops.mish_bwd(dgrad(q + "channel_mixer.2", dt), h1)
            wgrad(q + "channel_mixer.0", skip, dh1)
            g = dgrad(q + "channel_mixer.0", dh1, residual=g)             # + gradient of the block's skip
        wgrad("feats.0", S["xin"], g)

    def forward(self, x: Tensor) -> Tensor:
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if not need_grad:
            return self.engine_forward(x, save=False)[0]
        return _PlksrFn.apply(x, self, *self.parameters())


class _PlksrFn(torch.autograd.Function):
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


@ARCH_REGISTRY.register()
def realplksr_s(**kwargs):
    return realplksr(n_blocks=12, kernel_size=13, use_ea=False, **kwargs)
