"""tests/golden/otf_*.npz from the LIVE reference (build container only):  python -m oracle.make_golden_otf

otf_feed_data.npz — for a handful of seeds: the host decisions the reference's REAL `otf.feed_data` took
(recorded, oracle/ref_otf.py) and the LQ / GT batches it produced, over consecutive iterations so the
training-pair pool is exercised.  Inputs (GT, blur kernels) and the torch-RNG draws are regenerated from
the seed by the tests.  otf_kernels.npz — blur kernels of the reference's host synthesis for fixed seeds.
TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from __future__ import annotations

import json
import random
from pathlib import Path

import numpy as np
import torch

from oracle import ref_otf as R
from oracle import ref_shim

OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"
HOST_KEYS = ("scale1", "mode1", "gauss1", "blur2", "scale2", "mode2", "gauss2", "sinc_first", "mode3", "top", "left",
             "patch_size", "seed")
CASE = dict(batch=4, hr=128, scale=4, patch_size=24, queue_size=8, iters=8, seed0=100)


def case_inputs(seed: int, ds: dict, batch: int, hr: int):
    """GT batch + the three blur kernels per sample, from the seed (shared with the tests)."""
    from neosr_b200.data.degradations import synth_kernels
    gt = R.structured_gt(seed, batch, hr, hr)
    rng, pr = np.random.default_rng(seed), random.Random(seed)
    ks = [synth_kernels(ds, rng, pr) for _ in range(batch)]
    return (gt, *[torch.from_numpy(np.stack([k[i] for k in ks])) for i in range(3)])


def feed_data_cases():
    c = CASE
    ds = dict(R.DEGRADATIONS, patch_size=c["patch_size"], batch_size=c["batch"])
    out, model = {}, None
    for it in range(c["iters"]):
        seed = c["seed0"] + it
        gt, k1, k2, sk = case_inputs(seed, ds, c["batch"], c["hr"])
        lq_r, gt_r, plan, _, perm, model = R.run_reference(gt, k1, k2, sk, ds, c["scale"], seed, model=model,
                                                           queue_size=c["queue_size"])
        host = {k: (plan[k] if not isinstance(plan[k], (np.floating, np.integer)) else plan[k].item()) for k in HOST_KEYS}
        out[f"{it}.plan"] = np.frombuffer(json.dumps(host).encode(), dtype=np.uint8)
        out[f"{it}.lq"] = np.round(lq_r.numpy() * 255).astype(np.uint8)  # LQ lies on exact 8-bit levels (otf.py:251)
        assert np.array_equal(out[f"{it}.lq"].astype(np.float32) / np.float32(255), lq_r.numpy())
        out[f"{it}.gt"] = np.round(gt_r.numpy() * 255).astype(np.uint8)
        assert np.array_equal(out[f"{it}.gt"].astype(np.float32) / np.float32(255), gt_r.numpy())
        if perm is not None:
            out[f"{it}.perm"] = perm.numpy().astype(np.int16)
    np.savez_compressed(OUT / "otf_feed_data.npz", **out)


def kernel_cases():
    import neosr.data.degradations as D
    ds = R.DEGRADATIONS
    out = {}
    for seed in range(12):
        D.rng = np.random.default_rng(seed)
        random.seed(seed)
        out[f"mixed{seed}"] = D.random_mixed_kernels(ds["kernel_list"], ds["kernel_prob"], 7 + 2 * (seed % 8), ds["blur_sigma"],
                                                     ds["blur_sigma"], [-np.pi, np.pi], ds["betag_range"], ds["betap_range"],
                                                     noise_range=None).astype(np.float32)
        out[f"sinc{seed}"] = D.circular_lowpass_kernel(np.pi / 3 + 0.15 * seed, 7 + 2 * (seed % 8), pad_to=21).astype(np.float32)
    np.savez_compressed(OUT / "otf_kernels.npz", **out)

def loss_cases():
    """mssim_loss / consistency_loss values and input gradients from the reference modules (ssim_loss.py, consistency_loss.py)."""
    from neosr.losses.consistency_loss import consistency_loss
    from neosr.losses.ssim_loss import mssim_loss
    out = {}
    for tag, (x, gt) in loss_inputs().items():
        x = x.clone().requires_grad_(True)
        for name, mod in (("mssim", mssim_loss(loss_weight=1.0)), ("cons", consistency_loss(loss_weight=1.0)),
                          ("cons_noblur", consistency_loss(blur=False, saturation=1.1, brightness=0.95, loss_weight=0.5))):
            v = mod(x, gt)
            g, = torch.autograd.grad(v, x)
            out[f"{tag}.{name}.value"] = np.float32(v.item())
            out[f"{tag}.{name}.grad"] = g.numpy().copy()
    np.savez_compressed(OUT / "losses_ssim_consistency.npz", **out)


def loss_inputs():
    g = torch.Generator().manual_seed(11)
    gt = R.structured_gt(12, 2, 48, 40)
    far = (gt + 0.1 * torch.randn(gt.shape, generator=g)).clamp(-0.05, 1.05)
    near = gt + 0.002 * torch.randn(gt.shape, generator=g)  # cosim < 1e-3: the cosine branch is live
    return {"far": (far, gt), "near": (near, gt)}



HAT_TINY = dict(img_size=64, embed_dim=36, depths=(2, 2), num_heads=(3, 3), window_size=16, compress_ratio=3, squeeze_factor=6,
                conv_scale=0.01, overlap_ratio=0.5, mlp_ratio=2, upscale=4)


def hat_case():
    """Reference `hat` (hat_arch.py) on a tiny config: forward image, loss and a spread of parameter gradients."""
    from neosr.archs.hat_arch import hat

    from oracle.hat import HATConfig, hat_param_shapes
    from oracle.swinir import synth_params
    net = hat(drop_path_rate=0.0, upsampler="pixelshuffle", resi_connection="1conv", **HAT_TINY).train()
    p = synth_params(hat_param_shapes(HATConfig(**HAT_TINY)), seed=3)
    net.load_state_dict(p, strict=False)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 3, 32, 48, generator=g)
    y = net(x)
    gt = torch.rand(y.shape, generator=g)
    ((y - gt) ** 2).mean().backward()
    out = {"y": y.detach().numpy().copy()}
    for k, v in net.named_parameters():
        if k.startswith(("layers.0.residual_group.blocks.1.", "layers.1.residual_group.overlap_attn.", "conv_first", "conv_last", "norm.")):
            out["grad." + k] = v.grad.numpy().copy()
    np.savez_compressed(OUT / "hat_tiny_fwd_bwd.npz", **out)


def arch_cases():
    """compact / esrgan / realplksr / unet reference modules: forward output and the gradient norm of every parameter for
    (y**2).mean(), plus the U-Net's spectral-norm buffers after three training forwards; and the log + final parameter
    checksums of three REAL optimize_parameters iterations with the U-Net discriminator (C2's step shape)."""
    from oracle.compact import compact_param_shapes
    from oracle.esrgan import esrgan_param_shapes
    from oracle.realplksr import realplksr_param_shapes
    from oracle.swinir import SwinIRConfig, swinir_param_shapes, synth_params
    from oracle.unet import synth_unet
    out = {}

    def run(tag, net, p, x):
        net.load_state_dict(p)
        net.train()
        y = net(x)
        (y ** 2).mean().backward()
        out[f"{tag}.y"] = y.detach().numpy().copy()
        out[f"{tag}.gnorm"] = np.array([float(v.grad.double().norm()) for _, v in net.named_parameters()])

    g = torch.Generator().manual_seed(8)
    run("compact", ref_shim.build_network({"type": "compact", "upscale": 2, "num_conv": 4, "num_feat": 32}),
        synth_params(compact_param_shapes(num_feat=32, num_conv=4, upscale=2), seed=7), torch.rand(2, 3, 16, 24, generator=g))
    run("esrgan", ref_shim.build_network({"type": "esrgan", "scale": 4, "num_block": 2, "num_feat": 32, "num_grow_ch": 16}),
        synth_params(esrgan_param_shapes(scale=4, num_feat=32, num_block=2, num_grow_ch=16), seed=11), torch.rand(2, 3, 16, 24, generator=g))
    kw = dict(dim=32, n_blocks=2, upscaling_factor=4, kernel_size=17, use_ea=True)
    run("realplksr", ref_shim.build_network({"type": "realplksr", **kw}), synth_params(realplksr_param_shapes(**kw), seed=13),
        torch.rand(2, 3, 20, 24, generator=g))
    net = ref_shim.build_network({"type": "unet", "num_feat": 16})
    dp, db = synth_unet(num_feat=16, seed=21)
    net.load_state_dict({**dp, **db})
    net.train()
    for it in range(3):
        y = net(torch.rand(2, 3, 32, 40, generator=g))
    out["unet.y3"] = y.detach().numpy().copy()
    for k, v in net.named_buffers():
        out[f"unet.buf.{k}"] = v.numpy().copy()
    # GAN step (reference's real closure / optimize_parameters)
    from neosr.archs.swinir_arch import swinir
    from neosr.losses.basic_loss import L1Loss
    from neosr.losses.gan_loss import gan_loss

    from oracle.make_golden import OPTIM, TINY
    p = synth_params(swinir_param_shapes(SwinIRConfig(**TINY)), seed=4)
    netg = swinir(drop_path_rate=0.0, **TINY)
    netg.load_state_dict(p, strict=False)
    netd = ref_shim.build_network({"type": "unet", "num_feat": 16})
    dp, db = synth_unet(num_feat=16, seed=9)
    netd.load_state_dict({**dp, **db})
    model = ref_shim.make_image_model(netg.train(), cri_pix=L1Loss(1.0), optim_kw=OPTIM, net_d=netd, cri_gan=gan_loss("bce", loss_weight=0.3))
    gg = torch.Generator().manual_seed(6)
    for it in range(3):
        model.feed_data({"lq": torch.rand(2, 3, 16, 16, generator=gg), "gt": torch.rand(2, 3, 64, 64, generator=gg)})
        model.optimize_parameters(it)
        log = model.get_current_log()
        out[f"gan.log{it}"] = np.frombuffer(json.dumps({k: float(v) for k, v in log.items()}).encode(), dtype=np.uint8)
    out["gan.g_norms"] = np.array([float(v.detach().double().norm()) for _, v in netg.named_parameters()])
    out["gan.d_norms"] = np.array([float(v.detach().double().norm()) for _, v in netd.named_parameters()])
    np.savez_compressed(OUT / "archs_fwd.npz", **out)


if __name__ == "__main__":
    assert ref_shim.available(), "needs /root/reference"
    ref_shim.activate(4)
    feed_data_cases()
    kernel_cases()
    loss_cases()
    hat_case()
    arch_cases()
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)
