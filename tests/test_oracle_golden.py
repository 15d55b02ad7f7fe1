"""Oracle vs the committed golden fixtures (made by oracle/make_golden.py from the live
reference).  Runs everywhere on CPU; this is what pins the oracle on the GPU box."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import losses as OL
from oracle.make_golden import OPTIM, TINY
from oracle.step import make_swinir_trainer
from oracle.swinir import SwinIRConfig, swinir_forward, swinir_medium_config, swinir_param_shapes, synth_params

G = Path(__file__).parent / "golden"


def _rel(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("tag,hw", [("a", (16, 16)), ("b", (24, 32))])
def test_tiny_fwd_bwd(tag, hw):
    z = np.load(G / "swinir_tiny_fwd_bwd.npz")
    cfg = SwinIRConfig(**TINY)
    p = {k: v.requires_grad_(True) for k, v in synth_params(swinir_param_shapes(cfg), seed=1).items()}
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 3, *hw, generator=g)
    y = swinir_forward(p, cfg, x)
    gt = torch.rand(y.shape, generator=g)
    assert _rel(y.detach(), z[f"{tag}.y"]) < 1e-5
    grads = torch.autograd.grad((y - gt).abs().mean(), list(p.values()))
    for k, gi in zip(p, grads):
        assert _rel(gi, z[f"{tag}.grad.{k}"]) < 2e-4, k
    grads = torch.autograd.grad(((swinir_forward(p, cfg, x) - gt) ** 2).mean(), list(p.values()))
    for k, gi in zip(p, grads):
        assert _rel(gi, z[f"{tag}.mse_grad.{k}"]) < 2e-4, k


def test_tiny_step3():
    z = np.load(G / "swinir_tiny_step3.npz")
    cfg = SwinIRConfig(**TINY)
    p = synth_params(swinir_param_shapes(cfg), seed=4)
    vgg_p = synth_params(OL.vgg19_conv_shapes(), seed=5)
    tr = make_swinir_trainer(p, cfg, pixel_weight=1.0, percep_weight=0.5, vgg_params=vgg_p, optim=OPTIM, ema=0.999)
    g = torch.Generator().manual_seed(6)
    for it in range(3):
        lq, gt = torch.rand(2, 3, 16, 16, generator=g), torch.rand(2, 3, 64, 64, generator=g)
        tr.feed_data({"lq": lq, "gt": gt})
        tr.optimize_parameters(it)
        for k, v in tr.get_current_log().items():
            ref = float(z[f"log{it}.{k}"])
            assert abs(v - ref) <= 1e-5 * max(1.0, abs(ref)), (it, k)
    for i, k in enumerate(tr.names):
        assert _rel(tr.params[k].detach(), z[f"param.{k}"]) < 1e-4, k
        assert _rel(tr.ema.avg[i], z[f"ema.{k}"]) < 1e-4, k


def test_medium_forward_and_losses():
    z = np.load(G / "swinir_medium_fwd_loss.npz")
    cfg = swinir_medium_config(4)
    p = synth_params(swinir_param_shapes(cfg), seed=0)
    vgg_p = synth_params(OL.vgg19_conv_shapes(), seed=5)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 3, 64, 64, generator=g)
    gt = torch.rand(1, 3, 256, 256, generator=g)
    with torch.no_grad():
        y = swinir_forward(p, cfg, x)
        assert _rel(y, z["y"]) < 1e-5
        assert abs(float(OL.l1_loss(y, gt)) - float(z["l_g_pix"])) < 1e-6
        lp = float(OL.vgg_perceptual_loss(vgg_p, y, gt, 0.5))
        assert abs(lp - float(z["l_g_percep"])) < 1e-5 * max(1, abs(lp))
