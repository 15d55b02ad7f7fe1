"""Generate tests/golden/*.npz from the LIVE reference (build container only).

    python -m oracle.make_golden

Inputs and weights are regenerated from seeds by the tests (oracle.swinir.synth_params,
torch.Generator on CPU), so the fixtures only hold the reference's OUTPUTS: forward
images, parameter gradients, loss logs, updated parameters / EMA after k steps.
TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from oracle import losses as OL
from oracle import ref_shim
from oracle.swinir import SwinIRConfig, swinir_medium_config, swinir_param_shapes, synth_params

OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"

TINY = dict(img_size=16, embed_dim=36, depths=(2, 2), num_heads=(3, 3), window_size=8, mlp_ratio=2.0,
            upsampler="pixelshuffle", resi_connection="1conv", upscale=4)
OPTIM = dict(lr=1e-3, betas=(0.98, 0.92, 0.987), weight_decay=0.02, schedule_free=True, warmup_steps=1600)


def _ref_swinir(kw, params):
    from neosr.archs.swinir_arch import swinir
    net = swinir(drop_path_rate=0.0, **kw)
    net.load_state_dict(params, strict=False)
    return net.train()


def tiny_fwd_bwd():
    cfg = SwinIRConfig(**TINY)
    p = synth_params(swinir_param_shapes(cfg), seed=1)
    net = _ref_swinir(TINY, p)
    out = {}
    for tag, hw in (("a", (16, 16)), ("b", (24, 32))):
        g = torch.Generator().manual_seed(2)
        x = torch.rand(2, 3, *hw, generator=g)
        y = net(x)
        gt = torch.rand(y.shape, generator=g)
        net.zero_grad()
        (y - gt).abs().mean().backward()
        out[f"{tag}.y"] = y.detach().numpy()
        for k, v in net.named_parameters():
            out[f"{tag}.grad.{k}"] = v.grad.numpy().copy()
        # smooth loss: gradients are continuous in the forward value (no sign() flips), which is
        # what a 1e-3 relative bound on a split-precision engine can be held to
        net.zero_grad()
        ((net(x) - gt) ** 2).mean().backward()
        for k, v in net.named_parameters():
            out[f"{tag}.mse_grad.{k}"] = v.grad.numpy().copy()
    np.savez_compressed(OUT / "swinir_tiny_fwd_bwd.npz", **out)


def tiny_step():
    from neosr.losses.basic_loss import L1Loss
    cfg = SwinIRConfig(**TINY)
    p = synth_params(swinir_param_shapes(cfg), seed=4)
    vgg_p = synth_params(OL.vgg19_conv_shapes(), seed=5)
    net = _ref_swinir(TINY, p)
    cri_p = ref_shim.build_vgg_perceptual(vgg_p, loss_weight=0.5)
    model = ref_shim.make_image_model(net, cri_pix=L1Loss(1.0), cri_perceptual=cri_p, optim_kw=OPTIM)
    g = torch.Generator().manual_seed(6)
    out = {}
    for it in range(3):
        lq, gt = torch.rand(2, 3, 16, 16, generator=g), torch.rand(2, 3, 64, 64, generator=g)
        model.feed_data({"lq": lq, "gt": gt})
        model.optimize_parameters(it)
        for k, v in model.get_current_log().items():
            out[f"log{it}.{k}"] = np.float64(v)
    for k, v in net.named_parameters():
        out[f"param.{k}"] = v.detach().numpy().copy()
    for k, v in model.net_g_ema.module.named_parameters():
        out[f"ema.{k}"] = v.detach().numpy().copy()
    np.savez_compressed(OUT / "swinir_tiny_step3.npz", **out)


def medium_fwd_loss():
    cfg = swinir_medium_config(4)
    p = synth_params(swinir_param_shapes(cfg), seed=0)
    vgg_p = synth_params(OL.vgg19_conv_shapes(), seed=5)
    net = ref_shim.build_network({"type": "swinir_medium", "drop_path_rate": 0.0})
    net.load_state_dict(p, strict=False)
    net.train()
    from neosr.losses.basic_loss import L1Loss
    cri_p = ref_shim.build_vgg_perceptual(vgg_p, loss_weight=0.5)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 3, 64, 64, generator=g)
    gt = torch.rand(1, 3, 256, 256, generator=g)
    y = net(x)
    l_pix = L1Loss(1.0)(y, gt)
    l_per = cri_p(y, gt)
    (l_pix + l_per).backward()
    out = {"y": y.detach().numpy(), "l_g_pix": np.float64(l_pix.item()), "l_g_percep": np.float64(l_per.item())}
    keep = ("conv_first.weight", "layers.0.residual_group.blocks.0.attn.qkv.weight",
            "layers.0.residual_group.blocks.1.attn.relative_position_bias_table",
            "layers.2.residual_group.blocks.1.attn.relative_position_bias_table",
            "layers.5.residual_group.blocks.5.mlp.fc2.weight", "layers.3.conv.bias",
            "layers.1.residual_group.blocks.3.norm1.weight", "norm.bias",
            "conv_before_upsample.0.weight", "upsample.2.bias", "conv_last.weight")
    gn = 0.0
    for k, v in net.named_parameters():
        gn += float(v.grad.double().pow(2).sum())
        if k in keep:
            out[f"grad.{k}"] = v.grad.numpy().copy()
    out["grad_norm"] = np.float64(gn ** 0.5)
    np.savez_compressed(OUT / "swinir_medium_fwd_loss.npz", **out)


if __name__ == "__main__":
    assert ref_shim.available(), "needs /root/reference"
    ref_shim.activate(4)
    OUT.mkdir(parents=True, exist_ok=True)
    tiny_fwd_bwd()
    tiny_step()
    medium_fwd_loss()
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)
