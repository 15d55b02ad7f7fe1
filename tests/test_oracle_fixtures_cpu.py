"""OTF degradation pipeline — CPU side: the oracle against the committed fixtures of the reference's REAL
`otf.feed_data` (tests/golden/otf_feed_data.npz, made by oracle/make_golden_otf.py), the host blur-kernel
synthesis against the reference's kernels, and the plan drawing logic."""
import json
import random
from pathlib import Path

import numpy as np
import torch

from neosr_b200.data.degradations import circular_lowpass_kernel, random_mixed_kernels
from neosr_b200.models.otf import MODES, draw_plan
from oracle import otf as O
from oracle.make_golden_otf import CASE, case_inputs
from oracle.ref_otf import DEGRADATIONS

G = Path(__file__).parent / "golden"


def _plan(z, it):
    return json.loads(bytes(z[f"{it}.plan"]).decode())


def test_oracle_reproduces_reference_feed_data():
    """Bit-exact: inputs and torch-RNG draws are regenerated from the seed, the host decisions come from the
    fixture, the outputs must equal what the reference's own feed_data produced (pool included)."""
    z = np.load(G / "otf_feed_data.npz")
    c = CASE
    ds = dict(DEGRADATIONS, patch_size=c["patch_size"], batch_size=c["batch"])
    pool = O.Pool(c["queue_size"])
    dequeued = 0
    for it in range(c["iters"]):
        seed = c["seed0"] + it
        gt, k1, k2, sk = case_inputs(seed, ds, c["batch"], c["hr"])
        gen = torch.Generator().manual_seed(seed)
        lq, gtc, _, _ = O.degrade(gt, k1, k2, sk, _plan(z, it), c["scale"], ds=ds, gen=gen)
        perm = None
        if f"{it}.perm" in z:
            perm = torch.randperm(c["queue_size"], generator=gen)  # the reference draws it right after (otf.py:70)
            assert np.array_equal(perm.numpy(), z[f"{it}.perm"])
            dequeued += 1
        lq, gtc = pool.step(lq, gtc, perm)
        assert np.array_equal(np.round(lq.numpy() * 255).astype(np.uint8), z[f"{it}.lq"]), it
        assert np.array_equal(lq.numpy(), z[f"{it}.lq"].astype(np.float32) / np.float32(255)), it
        assert np.array_equal(gtc.numpy(), z[f"{it}.gt"].astype(np.float32) / np.float32(255)), it
    assert dequeued >= 4


def test_host_decisions_match_reference_for_equal_seeds():
    z = np.load(G / "otf_feed_data.npz")
    c = CASE
    ds = dict(DEGRADATIONS, patch_size=c["patch_size"], batch_size=c["batch"])
    for it in range(c["iters"]):
        seed = c["seed0"] + it
        p = draw_plan(ds, c["batch"], c["hr"], c["hr"], c["scale"], np.random.default_rng(seed), random.Random(seed),
                      np.random.default_rng(7))
        ref = _plan(z, it)
        for k in ("scale1", "mode1", "gauss1", "blur2", "scale2", "mode2", "gauss2", "sinc_first", "mode3", "top", "left"):
            assert p[k] == ref[k], (it, k, p[k], ref[k])


def test_plan_ranges():
    ds = dict(DEGRADATIONS, patch_size=24, batch_size=8)
    rng, pr, rd = np.random.default_rng(0), random.Random(0), np.random.default_rng(1)
    seen_modes, gauss = set(), 0
    for _ in range(300):
        p = draw_plan(ds, 8, 128, 128, 4, rng, pr, rd)
        assert p["scale1"] == 1 or 0.5 <= p["scale1"] <= 1.5
        assert p["scale2"] == 1 or 0.3 <= p["scale2"] <= 1.5
        assert {p["mode1"], p["mode2"], p["mode3"]} <= set(MODES)
        seen_modes |= {p["mode1"]}
        for i, sfx in ((1, ""), (2, "2")):
            if p[f"gauss{i}"]:
                gauss += 1
                lo, hi = ds["noise_range" + sfx]
                assert p[f"sigma{i}"].shape == (8,) and (p[f"sigma{i}"] >= lo).all() and (p[f"sigma{i}"] <= hi).all()
            else:
                lo, hi = ds["poisson_scale_range" + sfx]
                assert (p[f"pscale{i}"] >= lo).all() and (p[f"pscale{i}"] <= hi).all()
            assert set(np.unique(p[f"gray{i}"])) <= {0.0, 1.0}
        assert (p["jpeg_q1"] >= 40).all() and (p["jpeg_q1"] <= 95).all()
        assert (p["jpeg_q2"] >= 35).all() and (p["jpeg_q2"] <= 95).all()
        assert 0 <= p["top"] <= 32 - 24 and 0 <= p["left"] <= 32 - 24
    assert seen_modes == set(MODES)
    assert 60 < gauss < 180  # gaussian_noise_prob = 0.2 on 600 draws


def test_blur_kernels_match_reference():
    z = np.load(G / "otf_kernels.npz")
    ds = DEGRADATIONS
    for seed in range(12):
        k = 7 + 2 * (seed % 8)
        mine = random_mixed_kernels(ds["kernel_list"], ds["kernel_prob"], k, ds["blur_sigma"], ds["blur_sigma"],
                                    [-np.pi, np.pi], ds["betag_range"], ds["betap_range"], np.random.default_rng(seed),
                                    random.Random(seed))
        np.testing.assert_allclose(mine.astype(np.float32), z[f"mixed{seed}"], rtol=1e-6, atol=1e-9)
        assert abs(mine.sum() - 1) < 1e-12
        sinc = circular_lowpass_kernel(np.pi / 3 + 0.15 * seed, k, pad_to=21)
        assert sinc.shape == (21, 21)
        np.testing.assert_allclose(sinc.astype(np.float32), z[f"sinc{seed}"], rtol=1e-6, atol=1e-9)


def test_oracle_jpeg_fp64_agrees_with_fp32():
    """The fp64 twin of the oracle JPEG (tie-breaker for quantiser flips) agrees with fp32 except on flips."""
    x = torch.rand(2, 3, 40, 56, generator=torch.Generator().manual_seed(0))
    q = torch.tensor([30.0, 80.0])
    a, b = O.jpeg(x, q), O.jpeg(x.double(), q.double()).float()
    d = (a - b).abs()
    assert float((d > 1e-4).float().mean()) < 2e-3


def test_oracle_mssim_consistency_vs_reference_fixture():
    """oracle.losses.{msssim_loss, consistency_loss} (values + input gradients) against fixtures written from the
    reference modules; the `near` case has the data-dependent cosine branch active (consistency_loss.py:186-190)."""
    from oracle import losses as OL
    from oracle.make_golden_otf import loss_inputs
    z = np.load(G / "losses_ssim_consistency.npz")
    for tag, (x, gt) in loss_inputs().items():
        x = x.clone().requires_grad_(True)
        for name, fn in (("mssim", lambda a, b: OL.msssim_loss(a, b, 1.0)), ("cons", lambda a, b: OL.consistency_loss(a, b, 1.0)),
                         ("cons_noblur", lambda a, b: OL.consistency_loss(a, b, 0.5, blur=False, saturation=1.1, brightness=0.95))):
            v = fn(x, gt)
            g, = torch.autograd.grad(v, x)
            assert abs(float(v) - float(z[f"{tag}.{name}.value"])) <= 1e-6 * abs(float(v)), (tag, name)
            ref = torch.from_numpy(z[f"{tag}.{name}.grad"])
            assert float((g - ref).abs().max() / ref.abs().max()) < 1e-5, (tag, name)
    near = loss_inputs()["near"]
    assert float(OL.consistency_loss(*near)) != float(OL.consistency_loss(*near, force_cosim=False))


def test_oracle_hat_vs_reference_fixture():
    """oracle.hat.hat_forward + autograd against a fixture written from the reference `hat` module (tiny config, both HAB
    kinds, OCAB with its wrapping index, CAB): pins the HAT oracle wherever the reference tree is absent."""
    from oracle.hat import HATConfig, hat_forward, hat_param_shapes
    from oracle.make_golden_otf import HAT_TINY
    from oracle.swinir import synth_params
    z = np.load(G / "hat_tiny_fwd_bwd.npz")
    cfg = HATConfig(**HAT_TINY)
    p = {k: v.requires_grad_(True) for k, v in synth_params(hat_param_shapes(cfg), seed=3).items()}
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 3, 32, 48, generator=g)
    y = hat_forward(p, cfg, x)
    gt = torch.rand(y.shape, generator=g)
    ref_y = torch.from_numpy(z["y"])
    assert float((y.detach() - ref_y).abs().max() / ref_y.abs().max()) < 1e-5
    keys = [k[5:] for k in z.files if k.startswith("grad.")]
    assert len(keys) > 30
    grads = torch.autograd.grad(((y - gt) ** 2).mean(), [p[k] for k in keys])
    for k, gi in zip(keys, grads):
        r = torch.from_numpy(z["grad." + k])
        assert float((gi - r).abs().max() / r.abs().max().clamp_min(1e-30)) < 2e-4, k


def test_oracle_archs_and_gan_step_vs_reference_fixture():
    """compact / esrgan / realplksr / unet oracles and the oracle GAN step against fixtures written from the reference
    modules and its REAL optimize_parameters (tests/golden/archs_fwd.npz)."""
    from oracle.compact import compact_forward, compact_param_shapes
    from oracle.esrgan import esrgan_forward, esrgan_param_shapes
    from oracle.make_golden import OPTIM, TINY
    from oracle.realplksr import realplksr_forward, realplksr_param_shapes
    from oracle.step import make_swinir_trainer
    from oracle.swinir import SwinIRConfig, swinir_param_shapes, synth_params
    from oracle.unet import synth_unet, unet_forward
    z = np.load(G / "archs_fwd.npz")

    def check(tag, fn, shapes, seed, x):
        p = {k: v.requires_grad_(True) for k, v in synth_params(shapes, seed=seed).items()}
        y = fn(p, x)
        ref = torch.from_numpy(z[f"{tag}.y"])
        assert float((y.detach() - ref).abs().max() / ref.abs().max()) < 1e-5, tag
        gr = torch.autograd.grad((y ** 2).mean(), list(p.values()))
        gn = np.array([float(t.double().norm()) for t in gr])
        np.testing.assert_allclose(gn, z[f"{tag}.gnorm"], rtol=2e-4, atol=1e-12, err_msg=tag)

    g = torch.Generator().manual_seed(8)
    check("compact", lambda p, x: compact_forward(p, x, num_conv=4, upscale=2), compact_param_shapes(num_feat=32, num_conv=4, upscale=2),
          7, torch.rand(2, 3, 16, 24, generator=g))
    check("esrgan", lambda p, x: esrgan_forward(p, x, scale=4, num_block=2), esrgan_param_shapes(scale=4, num_feat=32, num_block=2, num_grow_ch=16),
          11, torch.rand(2, 3, 16, 24, generator=g))
    kw = dict(dim=32, n_blocks=2, upscaling_factor=4, kernel_size=17, use_ea=True)
    check("realplksr", lambda p, x: realplksr_forward(p, x, n_blocks=2, kernel_size=17, use_ea=True), realplksr_param_shapes(**kw),
          13, torch.rand(2, 3, 20, 24, generator=g))
    dp, db = synth_unet(num_feat=16, seed=21)
    bo = {k: v.clone() for k, v in db.items()}
    for it in range(3):
        y = unet_forward(dp, bo, torch.rand(2, 3, 32, 40, generator=g), training=True)
    ref = torch.from_numpy(z["unet.y3"])
    assert float((y - ref).abs().max() / ref.abs().max()) < 1e-5
    for k, v in bo.items():
        r = torch.from_numpy(z[f"unet.buf.{k}"])
        assert float((v - r).abs().max() / r.abs().max().clamp_min(1e-30)) < 1e-5, k
    p = synth_params(swinir_param_shapes(SwinIRConfig(**TINY)), seed=4)
    dp, db = synth_unet(num_feat=16, seed=9)
    tr = make_swinir_trainer(p, SwinIRConfig(**TINY), pixel_weight=1.0, optim=OPTIM, ema=0.999, disc=(dp, db), gan_weight=0.3)
    gg = torch.Generator().manual_seed(6)
    for it in range(3):
        tr.feed_data({"lq": torch.rand(2, 3, 16, 16, generator=gg), "gt": torch.rand(2, 3, 64, 64, generator=gg)})
        tr.optimize_parameters(it)
        ref_log = json.loads(bytes(z[f"gan.log{it}"]).decode())
        for k, v in tr.get_current_log().items():
            assert abs(v - ref_log[k]) <= 1e-5 * max(1.0, abs(ref_log[k])), (it, k)
    np.testing.assert_allclose([float(tr.params[k].detach().double().norm()) for k in tr.names], z["gan.g_norms"], rtol=1e-5)
    np.testing.assert_allclose([float(tr.d_params[k].detach().double().norm()) for k in tr.d_names], z["gan.d_norms"], rtol=1e-5)
