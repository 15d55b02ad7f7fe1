"""SRVGGNetCompact (`compact`) on the B200 kernels — drop-in for neosr/archs/compact_arch.py
(same constructor keywords and `body.N.*` state_dict keys).  Forward: conv+PReLU fused in the
contraction epilogue (pre-activation saved for the backward pass), last conv, pixel shuffle and the
nearest-upsampled input skip fused with the NHWC->NCHW output transpose."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import ops
from ..engine import ParamSet
from ..registry import ARCH_REGISTRY
from .arch_util import net_opt


@ARCH_REGISTRY.register()
class compact(nn.Module):
    def __init__(self, num_in_ch=3, num_out_ch=3, num_feat=64, num_conv=16, upscale=None, act_type="prelu", **kwargs):
        super().__init__()
        if upscale is None:
            upscale = net_opt()[0]
        if act_type not in ("relu", "prelu", "leakyrelu"):
            raise ValueError(f"act_type {act_type!r} not supported")
        self.num_in_ch, self.num_out_ch, self.num_feat = num_in_ch, num_out_ch, num_feat
        self.num_conv, self.upscale, self.act_type = num_conv, upscale, act_type

        def act():
            if act_type == "relu":
                return nn.ReLU(inplace=True)
            if act_type == "prelu":
                return nn.PReLU(num_parameters=num_feat)
            return nn.LeakyReLU(negative_slope=0.1, inplace=True)

        self.body = nn.ModuleList()
        self.body.append(nn.Conv2d(num_in_ch, num_feat, 3, 1, 1))
        self.body.append(act())
        for _ in range(num_conv):
            self.body.append(nn.Conv2d(num_feat, num_feat, 3, 1, 1))
            self.body.append(act())
        self.body.append(nn.Conv2d(num_feat, num_out_ch * upscale * upscale, 3, 1, 1))
        self.upsampler = nn.PixelShuffle(upscale)
        self._ps: ParamSet | None = None

    def param_set(self) -> ParamSet:
        if self._ps is None or any(self._ps._params[n] is not p for n, p in self.named_parameters()):
            self._ps = ParamSet(self)
        return self._ps

    def train(self, mode: bool = True):
        if self._ps is not None:
            self._ps.invalidate_packed()
        return super().train(mode)

    def _act_kw(self, ps, k):
        if self.act_type == "prelu":
            return {"act": "prelu", "prelu": ps.p(f"body.{2 * k + 1}.weight")}
        if self.act_type == "relu":
            return {"act": "relu"}
        return {"act": "lrelu", "act_slope": 0.1}

    def engine_forward(self, x: Tensor, save: bool):
        if not x.is_cuda:
            raise RuntimeError("neosr_b200.compact runs on CUDA (sm_100a) only; there is no CPU path")
        x = x.contiguous().float()
        ps = self.param_set()
        ps.pack_all()  # one launch re-packs every weight image after an optimizer step
        t = ops.nchw_to_nhwc_affine(x, None, None)
        S = []
        for k in range(self.num_conv + 1):
            y, pre = ops.conv_fprop(t, ps.pw(f"body.{2 * k}.weight", need_dgrad=k > 0), ps.p(f"body.{2 * k}.bias"),
                                    want_pre=True, **self._act_kw(ps, k))
            if save:
                S.append((t, pre))
            t = y
        last = 2 * (self.num_conv + 1)
        c = ops.conv_fprop(t, ps.pw(f"body.{last}.weight"), ps.p(f"body.{last}.bias"))
        out = ops.pixel_shuffle(c, self.upscale)
        y = ops.nhwc_to_nchw_add_nearest(out, x, self.upscale)
        return y, ((S, t) if save else None)

    def engine_backward(self, saved, dy: Tensor) -> None:
        S, t_last = saved
        ps = self.param_set()
        ps.ensure_grads(dy.device)

        def bwd(name, x_in, g, need_dx=True):
            ops.conv_wgrad(x_in, g, ps.g(name + ".weight"), ps.g(name + ".bias"), 3, 3)
            return ops.conv_fprop(g, ps.pw(name + ".weight"), None, dgrad=True) if need_dx else None

        g = ops.nchw_to_nhwc_affine(dy.contiguous().float(), None, None)
        g = ops.pixel_unshuffle(g, self.upscale)
        g = bwd(f"body.{2 * (self.num_conv + 1)}", t_last, g)
        for k in reversed(range(self.num_conv + 1)):
            t_in, pre = S[k]
            if self.act_type == "prelu":
                g = ops.prelu_bwd(g, pre, ps.p(f"body.{2 * k + 1}.weight"), ps.g(f"body.{2 * k + 1}.weight"))
            else:
                g = ops.actgrad_mul(g, pre, "relu" if self.act_type == "relu" else "lrelu", 0.1)
            g = bwd(f"body.{2 * k}", t_in, g, need_dx=k > 0)

    def forward(self, x: Tensor) -> Tensor:
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if not need_grad:
            return self.engine_forward(x, save=False)[0]
        return _CompactFn.apply(x, self, *self.parameters())


class _CompactFn(torch.autograd.Function):
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
