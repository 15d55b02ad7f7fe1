"""VGG19 feature extractor for the perceptual loss on the B200 kernels — drop-in for
neosr/archs/vgg_arch.py:76-199 (same constructor, `forward(x) -> {layer: tensor}`, state_dict
keys `vgg_net.convN_M.{weight,bias}`, buffers `mean`/`std` = 0.5 / 0.25).

The extractor is frozen (requires_grad=False, vgg_arch.py:157-164), so backward is dgrad only:
each conv's dgrad epilogue applies the ReLU mask of the tensor it differentiates and adds the
loss gradient of a tap in the same pass; max-pool backward is fused with the ReLU mask too.
"""
from __future__ import annotations

import os
import warnings
from pathlib import Path

import torch
from torch import Tensor, nn

from .. import ops
from ..engine import ParamSet
from ..registry import ARCH_REGISTRY

VGG_PRETRAIN_PATH = "experiments/pretrained_models/vgg19-dcbb9e9d.pth"
_STAGE_CH = (64, 128, 256, 512, 512)
_STAGE_CONVS = (2, 2, 4, 4, 4)


def _vgg19_names() -> list:
    names = []
    for s, n in enumerate(_STAGE_CONVS, start=1):
        for i in range(1, n + 1):
            names += [f"conv{s}_{i}", f"relu{s}_{i}"]
        names.append(f"pool{s}")
    return names


NAMES = {"vgg19": _vgg19_names()}


class _ConvHolder(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, 3, 3))
        self.bias = nn.Parameter(torch.empty(cout))


def _load_vgg19_weights(net: nn.Module, names_used: list, allow_random_init: bool) -> str:
    """Fill `vgg_net.*` from (1) the reference's local path, (2) torchvision's cached/downloaded
    VGG19_Weights.DEFAULT, else (3) seeded random init if explicitly allowed."""
    conv_names = [n for n in names_used if n.startswith("conv")]
    tv_index = {n: i for i, n in enumerate(NAMES["vgg19"])}
    sd = None
    src = ""
    if Path(VGG_PRETRAIN_PATH).exists():
        sd = torch.load(VGG_PRETRAIN_PATH, map_location="cpu", weights_only=True)
        src = VGG_PRETRAIN_PATH
    elif not allow_random_init:
        try:
            from torchvision.models import VGG19_Weights, vgg19  # noqa: PLC0415
            sd = vgg19(weights=VGG19_Weights.DEFAULT).state_dict()
            src = "torchvision VGG19_Weights.DEFAULT"
        except Exception as e:  # no network / no cache
            raise RuntimeError(
                "VGG19 weights unavailable (no local file, torchvision download failed). Place "
                f"{VGG_PRETRAIN_PATH} or pass allow_random_init=True / NSR_VGG_RANDOM_INIT=1 for "
                "synthetic-weight benchmarking.") from e
    if sd is not None:
        with torch.no_grad():
            for n in conv_names:
                conv = getattr(net, n)
                conv.weight.copy_(sd[f"features.{tv_index[n]}.weight"])
                conv.bias.copy_(sd[f"features.{tv_index[n]}.bias"])
        return src
    warnings.warn("VGGFeatureExtractor: using SEEDED RANDOM VGG19 weights (no pretrained weights available); "
                  "perceptual-loss values are synthetic.", stacklevel=3)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for n in conv_names:
            conv = getattr(net, n)
            fan_in = conv.weight.shape[1] * 9
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
            conv.bias.copy_(torch.randn(conv.bias.shape, generator=g) * 0.01)
    return "seeded-random"


@ARCH_REGISTRY.register()
class VGGFeatureExtractor(nn.Module):
    def __init__(self, layer_name_list, vgg_type: str = "vgg19", use_input_norm: bool = True,
                 range_norm: bool = False, requires_grad: bool = False, remove_pooling: bool = False,
                 pooling_stride: int = 2, allow_random_init: bool | None = None) -> None:
        super().__init__()
        if vgg_type != "vgg19" or requires_grad or remove_pooling or pooling_stride != 2:
            raise NotImplementedError("neosr_b200.VGGFeatureExtractor: frozen vgg19 with 2x2 pooling only "
                                      "(the perceptual-loss configuration of the reference)")
        self.layer_name_list = list(layer_name_list)
        self.use_input_norm, self.range_norm = use_input_norm, range_norm
        names = NAMES["vgg19"]
        max_idx = max(names.index(v) for v in self.layer_name_list)
        self.names = names[: max_idx + 1]
        self.vgg_net = nn.Module()
        cin = 3
        for n in self.names:
            if n.startswith("conv"):
                cout = _STAGE_CH[int(n[4]) - 1]
                setattr(self.vgg_net, n, _ConvHolder(cin, cout))
                cin = cout
        if allow_random_init is None:
            allow_random_init = os.environ.get("NSR_VGG_RANDOM_INIT", "0") == "1"
        self.weight_source = _load_vgg19_weights(self.vgg_net, self.names, allow_random_init)
        for p in self.parameters():
            p.requires_grad = False
        if use_input_norm:
            self.register_buffer("mean", torch.tensor([0.5, 0.5, 0.5]).view(1, 3, 1, 1))
            self.register_buffer("std", torch.tensor([0.25, 0.25, 0.25]).view(1, 3, 1, 1))
        self._ps: ParamSet | None = None
        self._aff: dict = {}

    def param_set(self) -> ParamSet:
        if self._ps is None:
            self._ps = ParamSet(self.vgg_net, trainable=False)
        return self._ps

    def _consts(self, device):
        c = self._aff.get(device)
        if c is None:
            if self.use_input_norm:
                mean, std = self.mean.flatten().to(device), self.std.flatten().to(device)
            else:
                mean, std = torch.zeros(3, device=device), torch.ones(3, device=device)
            a, b = 1.0 / std, -mean / std
            if self.range_norm:  # x = (x + 1) / 2 first
                a, b = a * 0.5, b + 0.5 / std
            c = {"scale": a.contiguous(), "shift": b.contiguous()}
            self._aff[device] = c
        return c

    def engine_forward(self, x: Tensor, save: bool):
        """x [B,3,H,W] -> ({tap: NHWC pre-ReLU feature}, saved)."""
        ps = self.param_set()
        k = self._consts(x.device)
        t = ops.nchw_to_nhwc_affine(x.contiguous().float(), k["scale"], k["shift"])
        taps, S = {}, []
        last = self.names[-1]
        for n in self.names:
            if n.startswith("conv"):
                pw, b = ps.pw(n + ".weight"), ps.p(n + ".bias")
                if n == last:
                    taps[n] = ops.conv_fprop(t, pw, b)
                    y = None
                elif n in self.layer_name_list:
                    y, pre = ops.conv_fprop(t, pw, b, act="relu", want_pre=True)
                    taps[n] = pre
                else:
                    y = ops.conv_fprop(t, pw, b, act="relu")
                if save:
                    S.append((n, y))
                t = y
            elif n.startswith("pool"):
                t = ops.maxpool2(t)
        return taps, (S if save else None)

    def engine_backward(self, S: list, dtaps: dict) -> Tensor:
        """Gradient w.r.t. the input image [B,3,H,W] given d(loss)/d(tap) (NHWC)."""
        ps = self.param_set()
        convs = [n for n in self.names if n.startswith("conv")]
        ys = dict(S)
        g = None  # gradient w.r.t. the pre-ReLU output of conv `n`
        for idx in reversed(range(len(convs))):
            n = convs[idx]
            gp = dtaps.get(n) if g is None else g
            if gp is None:
                raise RuntimeError("engine_backward: last layer must be a tap")
            if idx == 0:
                gin = ops.conv_fprop(gp, ps.pw(n + ".weight"), None, dgrad=True)
                k = self._consts(gin.device)
                return ops.nhwc_to_nchw_affine(gin, k["scale"], None)
            prev = convs[idx - 1]
            pooled = self.names[self.names.index(n) - 1].startswith("pool")
            if pooled:
                gpool = ops.conv_fprop(gp, ps.pw(n + ".weight"), None, dgrad=True)
                g = ops.maxpool2_relu_bwd(ys[prev], gpool, dtaps.get(prev))
            else:
                g = ops.conv_fprop(gp, ps.pw(n + ".weight"), None, dgrad=True, actgrad="relu", aux=ys[prev],
                                   residual=dtaps.get(prev))
        raise AssertionError

    def forward(self, x: Tensor) -> dict:
        """Reference-shaped output: {layer: NCHW tensor} (inference of features only; the
        training path uses engine_forward/engine_backward via vgg_perceptual_loss)."""
        taps, _ = self.engine_forward(x, save=False)
        return {k: ops.nhwc_to_nchw_affine(v, None, None) for k, v in taps.items()}
