"""RRDBNet (`esrgan`) on the B200 kernels — drop-in for neosr/archs/esrgan_arch.py (same constructor
keywords, `conv_first / body.N.rdbK.convJ / conv_body / conv_up1 / conv_up2 / conv_hr / conv_last`
state_dict keys).

Dense blocks never materialise `torch.cat` (esrgan_arch.py:109-116): every RDB owns one NHWC slab
[B,H,W, num_feat + 4*grow]; conv_k reads the first `num_feat + (k-1)*grow` channels in place
(x_ld = slab width) and writes its LeakyReLU output into the next channel slice.  conv5's epilogue
computes `x5 * 0.2 + x` straight into the first slice of the NEXT block's slab.  In the backward
pass the slab gradient is accumulated in place by the dgrad epilogues (residual = output buffer)."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import ops
from ..engine import ParamSet
from ..registry import ARCH_REGISTRY
from .arch_util import net_opt


class _RDB(nn.Module):
    def __init__(self, nf, gc):
        super().__init__()
        self.conv1 = nn.Conv2d(nf, gc, 3, 1, 1)
        self.conv2 = nn.Conv2d(nf + gc, gc, 3, 1, 1)
        self.conv3 = nn.Conv2d(nf + 2 * gc, gc, 3, 1, 1)
        self.conv4 = nn.Conv2d(nf + 3 * gc, gc, 3, 1, 1)
        self.conv5 = nn.Conv2d(nf + 4 * gc, nf, 3, 1, 1)
        for m in (self.conv1, self.conv2, self.conv3, self.conv4, self.conv5):  # default_init_weights(..., 0.1)
            nn.init.kaiming_normal_(m.weight)
            m.weight.data *= 0.1
            nn.init.constant_(m.bias, 0)


class _RRDB(nn.Module):
    def __init__(self, nf, gc):
        super().__init__()
        self.rdb1, self.rdb2, self.rdb3 = _RDB(nf, gc), _RDB(nf, gc), _RDB(nf, gc)


@ARCH_REGISTRY.register()
class esrgan(nn.Module):
    def __init__(self, num_in_ch=3, num_out_ch=3, scale=None, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        if scale is None:
            scale = net_opt()[0]
        if scale not in (1, 2, 4):
            raise ValueError("esrgan supports scale 1, 2 and 4")
        self.scale, self.nf, self.gc, self.num_block = scale, num_feat, num_grow_ch, num_block
        cin = num_in_ch * (4 if scale == 2 else (16 if scale == 1 else 1))
        self.conv_first = nn.Conv2d(cin, num_feat, 3, 1, 1)
        self.body = nn.Sequential(*[_RRDB(num_feat, num_grow_ch) for _ in range(num_block)])
        self.conv_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)
        self._ps: ParamSet | None = None
        self._fifth: dict = {}

    def param_set(self) -> ParamSet:
        if self._ps is None or any(self._ps._params[n] is not p for n, p in self.named_parameters()):
            self._ps = ParamSet(self)
        return self._ps

    def train(self, mode: bool = True):
        if self._ps is not None:
            self._ps.invalidate_packed()
        return super().train(mode)

    def _scale02(self, batch: int, device):
        key = (batch, device)
        if key not in self._fifth:
            self._fifth[key] = torch.full((batch,), 0.2, dtype=torch.float32, device=device)
        return self._fifth[key]

    # ------------------------------------------------------------------ forward
    def engine_forward(self, x: Tensor, save: bool):
        if not x.is_cuda:
            raise RuntimeError("neosr_b200.esrgan runs on CUDA (sm_100a) only; there is no CPU path")
        x = x.contiguous().float()
        ps = self.param_set()
        ps.pack_all()  # one launch re-packs every weight image after an optimizer step
        nf, gc = self.nf, self.gc
        ld = nf + 4 * gc
        xin = ops.nchw_to_nhwc_affine(x, None, None)
        if self.scale in (1, 2):
            xin = ops.pixel_unshuffle(xin, 2 if self.scale == 2 else 4)  # esrgan_arch.py:60-79 (same index map)
        B, H, W, _ = xin.shape
        s02 = self._scale02(B, x.device)

        def new_slab():
            return torch.empty((B, H, W, ld), dtype=torch.float32, device=x.device)

        def conv(name, src, **kw):
            return ops.conv_fprop(src, ps.pw(name + ".weight"), ps.p(name + ".bias"), **kw)

        cur = new_slab()
        conv("conv_first", xin, out=ops.Slab(cur, 0, nf))
        first = cur
        slabs = []
        for b in range(self.num_block):
            rin = cur
            for r in (1, 2, 3):
                pre = f"body.{b}.rdb{r}."
                for k in range(1, 5):
                    conv(pre + f"conv{k}", ops.Slab(cur, 0, nf + (k - 1) * gc), act="lrelu", act_slope=0.2,
                         out=ops.Slab(cur, nf + (k - 1) * gc, gc))
                nxt = new_slab()
                conv(pre + "conv5", ops.Slab(cur, 0, ld), row_scale=s02, residual=ops.Slab(cur, 0, nf),
                     out=ops.Slab(nxt, 0, nf))                       # x5 * 0.2 + x
                slabs.append(cur)
                cur = nxt
            ops.axpby2d(ops.Slab(cur, 0, nf), 0.2, ops.Slab(rin, 0, nf), 1.0, out=ops.Slab(cur, 0, nf))  # out*0.2 + x
        feat = conv("conv_body", ops.Slab(cur, 0, nf), residual=ops.Slab(first, 0, nf))
        u1 = ops.nearest_up2(feat)
        f1 = conv("conv_up1", u1, act="lrelu", act_slope=0.2)
        u2 = ops.nearest_up2(f1)
        f2 = conv("conv_up2", u2, act="lrelu", act_slope=0.2)
        f3 = conv("conv_hr", f2, act="lrelu", act_slope=0.2)
        out = conv("conv_last", f3)
        y = ops.nhwc_to_nchw_affine(out, None, None)
        S = {"xin": xin, "slabs": slabs, "last": cur, "tail": (u1, f1, u2, f2, f3)} if save else None
        return y, S

    # ------------------------------------------------------------------ backward
    def engine_backward(self, S: dict, dy: Tensor) -> None:
        ps = self.param_set()
        ps.ensure_grads(dy.device)
        nf, gc = self.nf, self.gc
        ld = nf + 4 * gc
        u1, f1, u2, f2, f3 = S["tail"]

        def bwd(name, x_in, g, need_dx=True, **epi):
            ops.conv_wgrad(x_in, g, ps.g(name + ".weight"), ps.g(name + ".bias"), 3, 3)
            return ops.conv_fprop(g, ps.pw(name + ".weight"), None, dgrad=True, **epi) if need_dx else None

        g = ops.nchw_to_nhwc_affine(dy.contiguous().float(), None, None)
        g = ops.actgrad_mul(bwd("conv_last", f3, g), f3, "lrelu", 0.2)  # 3-channel dgrad runs in conv_small.cu
        g = bwd("conv_hr", f2, g, actgrad="lrelu", actgrad_slope=0.2, aux=f2)
        g = ops.nearest_up2_bwd(bwd("conv_up2", u2, g))
        g = ops.actgrad_mul(g, f1, "lrelu", 0.2)
        g = ops.nearest_up2_bwd(bwd("conv_up1", u1, g))           # grad w.r.t. feat = conv_first + conv_body(...)
        dfirst = g
        g = bwd("conv_body", ops.Slab(S["last"], 0, nf), g)       # grad w.r.t. the last RRDB's output
        B, H, W, _ = g.shape
        s02 = self._scale02(B, g.device)
        idx = len(S["slabs"])
        for b in reversed(range(self.num_block)):
            g_rrdb = g                                             # out*0.2 + x: the skip takes g unchanged
            g = ops.axpby2d(g, 0.2, None, 0.0)
            for r in (3, 2, 1):
                idx -= 1
                slab = S["slabs"][idx]
                pre = f"body.{b}.rdb{r}."
                dS = torch.zeros((B, H, W, ld), dtype=torch.float32, device=g.device)
                ops.axpby2d(g, 1.0, None, 0.0, out=ops.Slab(dS, 0, nf))      # x5*0.2 + x: skip path
                dy5 = ops.axpby2d(g, 0.2, None, 0.0)
                ops.conv_wgrad(ops.Slab(slab, 0, ld), dy5, ps.g(pre + "conv5.weight"), ps.g(pre + "conv5.bias"), 3, 3)
                ops.conv_fprop(dy5, ps.pw(pre + "conv5.weight"), None, dgrad=True, residual=ops.Slab(dS, 0, ld),
                               out=ops.Slab(dS, 0, ld))
                for k in (4, 3, 2, 1):
                    c0, cin = nf + (k - 1) * gc, nf + (k - 1) * gc
                    dyk = ops.actgrad_mul2d(ops.Slab(dS, c0, gc), ops.Slab(slab, c0, gc), "lrelu", 0.2)
                    ops.conv_wgrad(ops.Slab(slab, 0, cin), dyk, ps.g(pre + f"conv{k}.weight"),
                                   ps.g(pre + f"conv{k}.bias"), 3, 3)
                    ops.conv_fprop(dyk, ps.pw(pre + f"conv{k}.weight"), None, dgrad=True,
                                   residual=ops.Slab(dS, 0, cin), out=ops.Slab(dS, 0, cin))
                g = ops.axpby2d(ops.Slab(dS, 0, nf), 1.0, None, 0.0)
            g = ops.axpby2d(g, 1.0, g_rrdb, 1.0)
        g = ops.axpby2d(g, 1.0, dfirst, 1.0)
        bwd("conv_first", S["xin"], g, need_dx=False)

    def forward(self, x: Tensor) -> Tensor:
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if not need_grad:
            return self.engine_forward(x, save=False)[0]
        return _EsrganFn.apply(x, self, *self.parameters())


class _EsrganFn(torch.autograd.Function):
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
