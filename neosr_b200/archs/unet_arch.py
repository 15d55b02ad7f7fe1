"""U-Net discriminator with spectral normalisation on the B200 kernels — drop-in for
neosr/archs/unet_arch.py:10-67 (same constructor keywords and state_dict keys: `convK.weight_orig`,
`convK.weight_u`, `convK.weight_v` for the eight spectral-normalised convolutions).

* The three 4x4 stride-2 convolutions run as 3x3 stride-1 convolutions over the pixel-unshuffled
  input (weights remapped by nsr_conv4x4s2_remap), so they use the tcgen05 implicit-GEMM kernels.
* Spectral norm follows torch.nn.utils.spectral_norm: ONE power iteration per training-mode forward,
  updating the u / v buffers in place (the reference runs three discriminator forwards per GAN step,
  so the buffers advance three times), none in eval mode.  The backward of a forward must run before
  the next forward of the same module: sigma/u/v/W_sn are per-forward state kept in one place.
* `engine_backward` can (a) skip parameter gradients (generator pass: the reference freezes net_d),
  (b) accumulate into the gradient buffer (second discriminator pass), (c) skip the input gradient."""
from __future__ import annotations

import torch
from torch import Tensor, nn
from torch.nn.utils import spectral_norm

from .. import ops
from ..engine import ParamSet
from ..registry import ARCH_REGISTRY

_SN = {"conv1": 4, "conv2": 4, "conv3": 4, "conv4": 3, "conv5": 3, "conv6": 3, "conv7": 3, "conv8": 3}


@ARCH_REGISTRY.register()
class unet(nn.Module):
    def __init__(self, num_in_ch=3, num_feat=64, skip_connection=True):
        super().__init__()
        self.skip_connection = skip_connection
        nf = num_feat
        # construction mirrors the reference so initial weights / RNG consumption are identical;
        # torch's forward pre-hooks are never used: the engine below computes W / sigma itself.
        self.conv0 = nn.Conv2d(num_in_ch, nf, 3, 1, 1)
        self.conv1 = spectral_norm(nn.Conv2d(nf, nf * 2, 4, 2, 1, bias=False))
        self.conv2 = spectral_norm(nn.Conv2d(nf * 2, nf * 4, 4, 2, 1, bias=False))
        self.conv3 = spectral_norm(nn.Conv2d(nf * 4, nf * 8, 4, 2, 1, bias=False))
        self.conv4 = spectral_norm(nn.Conv2d(nf * 8, nf * 4, 3, 1, 1, bias=False))
        self.conv5 = spectral_norm(nn.Conv2d(nf * 4, nf * 2, 3, 1, 1, bias=False))
        self.conv6 = spectral_norm(nn.Conv2d(nf * 2, nf, 3, 1, 1, bias=False))
        self.conv7 = spectral_norm(nn.Conv2d(nf, nf, 3, 1, 1, bias=False))
        self.conv8 = spectral_norm(nn.Conv2d(nf, nf, 3, 1, 1, bias=False))
        self.conv9 = nn.Conv2d(nf, 1, 3, 1, 1)
        self._ps: ParamSet | None = None
        self._sn_state: dict = {}

    def param_set(self) -> ParamSet:
        if self._ps is None or any(self._ps._params[n] is not p for n, p in self.named_parameters()):
            self._ps = ParamSet(self)
        return self._ps

    def train(self, mode: bool = True):
        if self._ps is not None:
            self._ps.invalidate_packed()
        return super().train(mode)

    # ------------------------------------------------------------------ spectral norm
    def _sn(self, name: str):
        conv = getattr(self, name)
        w = conv.weight_orig.detach()
        st = self._sn_state.get(name)
        if st is None or st["w_sn"].device != w.device:
            cout, cin, k, _ = w.shape
            st = {"w_sn": torch.empty_like(w), "sigma": torch.empty(1, dtype=torch.float32, device=w.device)}
            eff = st["w_sn"]
            if k == 4:
                eff = st["w3"] = torch.empty((cout, 4 * cin, 3, 3), dtype=torch.float32, device=w.device)
            st["pk"] = ops.PackedWeight(eff)
            self._sn_state[name] = st
        return conv, w, st

    def _refresh_weights(self) -> None:
        """W_sn = W / sigma for the eight SN convolutions (unet_arch.py:28-38), then re-pack."""
        iters = 1 if self.training else 0
        for name, k in _SN.items():
            conv, w, st = self._sn(name)
            ops.spectral_norm_fwd(w, conv.weight_u, conv.weight_v, st["w_sn"], st["sigma"], iters, 1e-12)
            if k == 4:
                ops.conv4x4s2_remap(st["w_sn"], w.shape[0], w.shape[1], out=st["w3"])
            st["pk"].refresh(force=True)

    # ------------------------------------------------------------------ forward
    def engine_forward(self, x: Tensor, save: bool):
        if not x.is_cuda:
            raise RuntimeError("neosr_b200.unet runs on CUDA (sm_100a) only; there is no CPU path")
        B, _, H, W = x.shape
        if H % 8 or W % 8:
            raise ValueError("unet: input height/width must be multiples of 8")
        ps = self.param_set()
        self._refresh_weights()
        pk = {n: self._sn_state[n]["pk"] for n in _SN}
        skip = self.skip_connection
        lr = dict(act="lrelu", act_slope=0.2)
        xin = ops.nchw_to_nhwc_affine(x.contiguous().float(), None, None)
        x0 = ops.conv_fprop(xin, ps.pw("conv0.weight"), ps.p("conv0.bias"), **lr)
        x0u = ops.pixel_unshuffle(x0, 2)
        x1 = ops.conv_fprop(x0u, pk["conv1"], None, **lr)
        x1u = ops.pixel_unshuffle(x1, 2)
        x2 = ops.conv_fprop(x1u, pk["conv2"], None, **lr)
        x2u = ops.pixel_unshuffle(x2, 2)
        x3 = ops.conv_fprop(x2u, pk["conv3"], None, **lr)
        x3b = ops.bilinear_up2(x3)
        if skip:   # lrelu(conv(.)) + skip in the epilogue; the pre-activation is kept for lrelu'
            x4, p4 = ops.conv_fprop(x3b, pk["conv4"], None, residual=x2, want_pre=True, **lr)
        else:
            x4, p4 = ops.conv_fprop(x3b, pk["conv4"], None, **lr), None
        x4b = ops.bilinear_up2(x4)
        if skip:
            x5, p5 = ops.conv_fprop(x4b, pk["conv5"], None, residual=x1, want_pre=True, **lr)
        else:
            x5, p5 = ops.conv_fprop(x4b, pk["conv5"], None, **lr), None
        x5b = ops.bilinear_up2(x5)
        if skip:
            x6, p6 = ops.conv_fprop(x5b, pk["conv6"], None, residual=x0, want_pre=True, **lr)
        else:
            x6, p6 = ops.conv_fprop(x5b, pk["conv6"], None, **lr), None
        o7 = ops.conv_fprop(x6, pk["conv7"], None, **lr)
        o8 = ops.conv_fprop(o7, pk["conv8"], None, **lr)
        out = ops.conv_fprop(o8, ps.pw("conv9.weight"), ps.p("conv9.bias"))
        y = out.view(B, 1, H, W)  # one channel: NHWC and NCHW coincide
        S = None
        if save:
            S = {"xin": xin, "x0": x0, "x0u": x0u, "x1": x1, "x1u": x1u, "x2": x2, "x2u": x2u, "x3": x3, "x3b": x3b,
                 "a4": p4 if skip else x4, "x4b": x4b, "a5": p5 if skip else x5, "x5b": x5b, "a6": p6 if skip else x6,
                 "x6": x6, "o7": o7, "o8": o8}
        return y, S

    # ------------------------------------------------------------------ backward
    def engine_backward(self, S: dict, dy: Tensor, param_grads: bool = True, accumulate: bool = False,
                        need_dx: bool = True):
        ps = self.param_set()
        ps.ensure_grads(dy.device)
        skip = self.skip_connection
        dev = dy.device
        B, _, H, W = dy.shape

        def plain_wgrad(name, x_in, g):
            if not param_grads:
                return
            gw, gb = ps.g(name + ".weight"), ps.g(name + ".bias")
            if accumulate:
                tw, tb = torch.empty_like(gw), torch.empty_like(gb)
                ops.conv_wgrad(x_in, g, tw, tb, 3, 3)
                ops.axpby(gw, 1.0, tw, 1.0, out=gw)
                ops.axpby(gb, 1.0, tb, 1.0, out=gb)
            else:
                ops.conv_wgrad(x_in, g, gw, gb, 3, 3)

        def sn_wgrad(name, x_in, g):
            if not param_grads:
                return
            conv, w, st = self._sn(name)
            cout, cin, k, _ = w.shape
            if k == 4:
                g3 = torch.empty((cout, 4 * cin, 3, 3), dtype=torch.float32, device=dev)
                ops.conv_wgrad(x_in, g, g3, None, 3, 3)
                gsn = ops.conv4x4s2_remap(g3, cout, cin, inverse=True)
            else:
                gsn = torch.empty_like(w)
                ops.conv_wgrad(x_in, g, gsn, None, 3, 3)
            ops.spectral_norm_bwd(gsn, st["w_sn"], conv.weight_u, conv.weight_v, st["sigma"],
                                  ps.g(name + ".weight_orig"), accumulate)

        def dgrad(name, g, **epi):
            return ops.conv_fprop(g, self._sn_state[name]["pk"], None, dgrad=True, **epi)

        g = dy.contiguous().float().view(B, H, W, 1)
        plain_wgrad("conv9", S["o8"], g)
        g = ops.actgrad_mul(ops.conv_fprop(g, ps.pw("conv9.weight"), None, dgrad=True), S["o8"], "lrelu", 0.2)
        sn_wgrad("conv8", S["o7"], g)
        g = dgrad("conv8", g, actgrad="lrelu", actgrad_slope=0.2, aux=S["o7"])
        sn_wgrad("conv7", S["x6"], g)
        # grad w.r.t. x6 = lrelu(h6) + x0: y_pre keeps the raw gradient for the skip, y takes lrelu'(h6)
        gh, gskip0 = self._split(dgrad, "conv7", g, S["a6"], skip)
        sn_wgrad("conv6", S["x5b"], gh)
        g = ops.bilinear_up2_bwd(dgrad("conv6", gh))
        gskip1 = g if skip else None
        gh = ops.actgrad_mul(g, S["a5"], "lrelu", 0.2)
        sn_wgrad("conv5", S["x4b"], gh)
        g = ops.bilinear_up2_bwd(dgrad("conv5", gh))
        gskip2 = g if skip else None
        gh = ops.actgrad_mul(g, S["a4"], "lrelu", 0.2)
        sn_wgrad("conv4", S["x3b"], gh)
        g = ops.bilinear_up2_bwd(dgrad("conv4", gh))
        gh = ops.actgrad_mul(g, S["x3"], "lrelu", 0.2)
        sn_wgrad("conv3", S["x2u"], gh)
        gh = self._down(dgrad("conv3", gh), gskip2, S["x2"])
        sn_wgrad("conv2", S["x1u"], gh)
        gh = self._down(dgrad("conv2", gh), gskip1, S["x1"])
        sn_wgrad("conv1", S["x0u"], gh)
        gh = self._down(dgrad("conv1", gh), gskip0, S["x0"])
        plain_wgrad("conv0", S["xin"], gh)
        if not need_dx:
            return None
        dx = ops.conv_fprop(gh, ps.pw("conv0.weight"), None, dgrad=True)
        return ops.nhwc_to_nchw_affine(dx, None, None)

    @staticmethod
    def _split(dgrad, name, g, pre, skip):
        if skip:
            gh, graw = dgrad(name, g, actgrad="lrelu", actgrad_slope=0.2, aux=pre, want_pre=True)
            return gh, graw
        return dgrad(name, g, actgrad="lrelu", actgrad_slope=0.2, aux=pre), None

    @staticmethod
    def _down(gu, gskip, act_out):
        """Gradient of one stride-2 stage: undo the space-to-depth view, add the skip branch's gradient,
        multiply by lrelu'(stage output)."""
        g = ops.pixel_shuffle(gu, 2)
        if gskip is not None:
            ops.axpby(g, 1.0, gskip, 1.0, out=g)
        return ops.actgrad_mul(g, act_out, "lrelu", 0.2)

    def forward(self, x: Tensor) -> Tensor:
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not need_grad:
            return self.engine_forward(x, save=False)[0]
        return _UnetFn.apply(x, self, *self.parameters())


class _UnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, net, *params):
        y, saved = net.engine_forward(x, save=True)
        ctx.net, ctx.saved = net, saved
        ctx.need_dx = x.requires_grad
        ctx.param_grads = any(p.requires_grad for p in params)
        return y

    @staticmethod
    def backward(ctx, dy):
        net = ctx.net
        dx = net.engine_backward(ctx.saved, dy, param_grads=ctx.param_grads, accumulate=False, need_dx=ctx.need_dx)
        ctx.saved = None
        ps = net.param_set()
        grads = [ps.g(n) if (p.requires_grad and ctx.param_grads) else None for n, p in net.named_parameters()]
        return (dx, None, *grads)
