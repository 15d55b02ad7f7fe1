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


if __name__ == "__main__":
    assert ref_shim.available(), "needs /root/reference"
    ref_shim.activate(4)
    feed_data_cases()
    kernel_cases()
    for f in sorted(OUT.glob("otf_*.npz")):
        print(f.name, f.stat().st_size)
