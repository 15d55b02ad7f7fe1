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



if __name__ == "__main__":
    assert ref_shim.available(), "needs /root/reference"
    ref_shim.activate(4)
    feed_data_cases()
    kernel_cases()
    loss_cases()
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)
