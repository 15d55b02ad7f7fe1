"""OTF degradation kernels (csrc/otf.cu) through the C ABI against the CPU oracle (oracle/otf.py, pinned to
the reference's own functions and its real otf.feed_data).  Tolerances: fp32 stencil / resample sums 2e-6
absolute on [0,1] images; exact where the arithmetic is order-free (noise application with given fields,
crop, quantise, pool); JPEG and the whole pipeline tolerate isolated quantiser flips (a DCT coefficient or
an 8-bit level within float rounding of a .5 boundary rounds the other way — the reference's own CPU and
CUDA paths differ the same way), bounded as a fraction of pixels."""
import json
import random
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from neosr_b200 import ops  # noqa: E402
from oracle import otf as O  # noqa: E402
from oracle.make_golden_otf import CASE, case_inputs  # noqa: E402
from oracle.ref_otf import DEGRADATIONS, structured_gt  # noqa: E402

G = Path(__file__).parent / "golden"
DEV = "cuda"


def _g(seed=0):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("k,kb", [(21, 3), (7, 3), (21, 1), (1, 3), (13, 1)])
@pytest.mark.parametrize("hw", [(50, 70), (128, 128), (33, 200)])
def test_filter2d(k, kb, hw):
    img = structured_gt(1, 3, *hw)
    ker = torch.rand(kb, k, k, generator=_g(k))
    ker /= ker.sum((1, 2), keepdim=True)
    ref = O.filter2d(img, ker)
    out = ops.filter2d(img.to(DEV), ker.to(DEV)).cpu()
    assert float((out - ref).abs().max()) < 2e-6


def test_filter2d_errors():
    img = torch.rand(2, 3, 32, 32, device=DEV)
    with pytest.raises(ValueError, match="Wrong kernel size"):
        ops.filter2d(img, torch.rand(1, 8, 8, device=DEV))
    with pytest.raises(RuntimeError):  # reflect padding needs an image larger than the pad
        ops.filter2d(torch.rand(1, 3, 8, 8, device=DEV), torch.rand(1, 21, 21, device=DEV))
    with pytest.raises(RuntimeError):
        ops.filter2d(img, torch.rand(3, 7, 7, device=DEV))  # kernel batch neither 1 nor B


@pytest.mark.parametrize("mode", ["area", "bilinear", "bicubic"])
@pytest.mark.parametrize("sf", [0.5, 0.6180339, 1.0, 1.4174908152510044, 0.31, 1.5])
def test_resize_scale_factor(mode, sf):
    img = structured_gt(2, 2, 61, 83)
    ref = O.resize(img, mode, scale_factor=sf)
    out = ops.resize(img.to(DEV), mode, scale_factor=sf).cpu()
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) < 2e-6


@pytest.mark.parametrize("mode", ["area", "bilinear", "bicubic"])
@pytest.mark.parametrize("size", [(32, 32), (9, 17), (100, 64), (61, 83), (1, 1)])
def test_resize_size(mode, size):
    img = structured_gt(3, 2, 61, 83)
    ref = O.resize(img, mode, size=size)
    out = ops.resize(img.to(DEV), mode, size=size).cpu()
    assert float((out - ref).abs().max()) < 2e-6


@pytest.mark.parametrize("gray", [[0.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 1.0, 1.0]])
def test_gaussian_noise_given_fields_is_exact(gray):
    g = _g(5)
    img = structured_gt(4, 3, 40, 56)
    z, zg = torch.randn(3, 3, 40, 56, generator=g), torch.randn(40, 56, generator=g)
    sigma, gray = torch.tensor([1.0, 10.0, 30.0]), torch.tensor(gray)
    ref = O.gaussian_noise(img, sigma, gray, z, zg)
    out = ops.gaussian_noise(img.to(DEV), sigma.to(DEV), gray.to(DEV), bool(gray.sum() > 0), 0, z.to(DEV), zg.to(DEV)).cpu()
    assert torch.equal(out, ref)


@pytest.mark.parametrize("gray", [[0.0, 0.0, 0.0], [0.0, 1.0, 1.0]])
def test_poisson_noise_given_counts_is_exact(gray):
    img = structured_gt(6, 3, 48, 40) * 0.8 + 0.1 * torch.rand(3, 3, 48, 40, generator=_g(1))
    img[2] = torch.round(img[2] * 7) / 7  # a sample with only 8 distinct levels -> vals = 8
    scale, gray = torch.tensor([0.05, 0.25, 1.0]), torch.tensor(gray)
    ref, cc, cg = O.poisson_noise(img, scale, gray, generator=_g(2))
    out = ops.poisson_noise(img.to(DEV), scale.to(DEV), gray.to(DEV), bool(gray.sum() > 0), 0, cc.to(DEV),
                            None if cg is None else cg.to(DEV)).cpu()
    assert torch.equal(out, ref)


def test_gaussian_noise_in_kernel_rng_statistics():
    """Philox + Box-Muller draws: N(0, (sigma/255)^2) per sample; the gray field is ONE [h,w] field shared by
    the batch (degradations.py:593-598) and identical across channels; seeds decorrelate."""
    B, H, W = 4, 192, 192
    img = torch.full((B, 3, H, W), 0.5, device=DEV)
    sigma = torch.tensor([2.0, 10.0, 25.0, 10.0], device=DEV)
    gray = torch.tensor([0.0, 0.0, 1.0, 1.0], device=DEV)
    out = ops.gaussian_noise(img, sigma, gray, True, 1234)
    n = (out - 0.5) * 255
    for b in range(B):
        x = n[b].flatten().double()
        s = float(sigma[b])
        assert abs(float(x.mean())) < 4 * s / (x.numel() ** 0.5) + 1e-3
        assert abs(float(x.std()) / s - 1) < 0.02
        kurt = float(((x - x.mean()) ** 4).mean() / x.var() ** 2)
        assert abs(kurt - 3) < 0.1
    assert torch.equal(n[2, 0], n[2, 1]) and torch.equal(n[2, 1], n[2, 2])      # gray: same over channels
    assert float((n[2, 0] / 25.0 - n[3, 0] / 10.0).abs().max()) < 1e-4          # and shared over the batch
    c01 = float(torch.corrcoef(torch.stack([n[0, 0].flatten(), n[0, 1].flatten()]))[0, 1])
    assert abs(c01) < 0.02
    out2 = ops.gaussian_noise(img, sigma, gray, True, 1235)
    c = float(torch.corrcoef(torch.stack([n[1].flatten(), ((out2 - 0.5) * 255)[1].flatten()]))[0, 1])
    assert abs(c) < 0.02
    assert torch.equal(out, ops.gaussian_noise(img, sigma, gray, True, 1234))    # deterministic per seed


def test_poisson_noise_in_kernel_rng_statistics():
    """Poisson(lambda) draws: mean = var = lambda over the whole lambda range both samplers cover."""
    B, H, W = 3, 256, 256
    levels = torch.arange(256, dtype=torch.float32).repeat_interleave(256).view(1, 1, H, W) / 255
    img = levels.expand(B, 3, H, W).contiguous().to(DEV)  # 256 distinct levels -> vals = 256, lambda = level/255*256
    scale, gray = torch.ones(B, device=DEV), torch.zeros(B, device=DEV)
    out = torch.stack([ops.poisson_noise(img, scale * 0 + 1e-3, gray, False, 77 + i) for i in range(4)])
    cnt = ((out - img) / 1e-3 + img) * 256  # recover counts: out = img + (cnt/256 - img) * 1e-3 (unclamped interior)
    for lvl in (1, 3, 9, 10, 40, 128, 200, 254):
        lam = lvl / 255 * 256
        x = cnt[:, :, :, lvl, :].flatten().double()
        assert abs(float(x.mean()) - lam) < 0.05 * max(1.0, lam ** 0.5), (lvl, float(x.mean()), lam)
        assert abs(float(x.var()) / lam - 1) < 0.08, (lvl, float(x.var()), lam)


@pytest.mark.parametrize("hw", [(64, 64), (50, 70), (16, 16), (37, 129)])
def test_jpeg(hw):
    img = structured_gt(7, 4, *hw)
    q = torch.tensor([35.0, 49.99, 50.0, 93.0])
    ref = O.jpeg(img, q)
    ref64 = O.jpeg(img.double(), q.double()).float()
    out = ops.jpeg(img.to(DEV), q.to(DEV)).cpu()
    d = torch.minimum((out - ref).abs(), (out - ref64).abs())
    assert float((d > 1e-5).float().mean()) < 2e-3      # isolated quantiser flips only
    assert float(d.median()) < 1e-6


def test_jpeg_clamps_input_and_respects_quality():
    img = structured_gt(8, 2, 48, 48) * 1.4 - 0.2
    q = torch.tensor([40.0, 90.0])
    ref = O.jpeg(img.clamp(0, 1), q)
    out = ops.jpeg(img.to(DEV), q.to(DEV)).cpu()
    assert float(((out - ref).abs() > 1e-5).float().mean()) < 2e-3
    clean = img.clamp(0, 1)
    assert float((out[0] - clean[0]).abs().mean()) > float((out[1] - clean[1]).abs().mean())


def test_crop_quantise_and_pool_are_exact():
    g = _g(3)
    x = torch.rand(4, 3, 40, 52, generator=g) * 1.2 - 0.1
    out = ops.crop(x.to(DEV), 5, 7, 24, 32, quantise=True).cpu()
    ref = torch.clamp((x[:, :, 5:29, 7:39] * 255.0).round(), 0, 255) / 255.0
    assert torch.equal(out, ref)
    assert torch.equal(ops.crop(x.to(DEV), 16, 20, 24, 32).cpu(), x[:, :, 16:40, 20:52])
    with pytest.raises(RuntimeError):
        ops.crop(x.to(DEV), 20, 0, 24, 32)
    pool = torch.rand(8, 3, 6, 6, generator=g)
    new = torch.rand(2, 3, 6, 6, generator=g)
    dpool = pool.to(DEV)
    slots = torch.tensor([5, 2], dtype=torch.int32, device=DEV)
    got = ops.pool_swap(dpool, new.to(DEV), slots, dequeue=True).cpu()
    assert torch.equal(got, pool[[5, 2]])
    exp = pool.clone()
    exp[[5, 2]] = new
    assert torch.equal(dpool.cpu(), exp)
    assert ops.pool_swap(dpool, new.to(DEV) * 0, slots, dequeue=False) is None
    exp[[5, 2]] = 0
    assert torch.equal(dpool.cpu(), exp)


def _otf_model(ds, queue_size):
    from neosr_b200.models.otf import otf
    m = object.__new__(otf)  # pipeline-only instance: no generator network needed for feed_data's kernels
    m.opt = {"scale": CASE["scale"], "datasets": {"train": ds}}
    m.device = torch.device("cuda", torch.cuda.current_device())
    m._ds, m.queue_size, m.queue_ptr, m.queue_lr, m.queue_gt = ds, queue_size, 0, None, None
    m._perm = np.arange(queue_size)
    m._rng_dev = np.random.default_rng(0)
    return m


def test_pipeline_replay_vs_oracle_and_reference_fixture():
    """Whole feed_data pipeline + pool over 8 iterations on the decisions the REFERENCE took (fixture) and
    the random fields the oracle drew: LQ must sit on 8-bit levels and equal the oracle's (== the
    reference's, tests/test_oracle_fixtures_cpu.py) except for isolated quantiser flips; GT crops are exact."""
    z = np.load(G / "otf_feed_data.npz")
    c = CASE
    ds = dict(DEGRADATIONS, patch_size=c["patch_size"], batch_size=c["batch"])
    pool = O.Pool(c["queue_size"])
    m = _otf_model(ds, c["queue_size"])
    worst = 0.0
    for it in range(c["iters"]):
        seed = c["seed0"] + it
        gt, k1, k2, sk = case_inputs(seed, ds, c["batch"], c["hr"])
        gen = torch.Generator().manual_seed(seed)
        host_plan = json.loads(bytes(z[f"{it}.plan"]).decode())
        lq_o, gt_o, plan, fields = O.degrade(gt, k1, k2, sk, host_plan, c["scale"], ds=ds, gen=gen)
        perm = torch.randperm(c["queue_size"], generator=gen) if f"{it}.perm" in z else None
        lq_o, gt_o = pool.step(lq_o, gt_o, perm)
        dev_fields = {k: v.to(DEV).contiguous() for k, v in fields.items()}
        lq, gtc = m.run_plan(gt.to(DEV), k1.to(DEV), k2.to(DEV), sk.to(DEV), plan, dev_fields)
        lq, gtc = m._dequeue_and_enqueue(lq, gtc, None if perm is None else perm.numpy())
        lq, gtc = lq.cpu(), gtc.cpu()
        assert torch.equal(gtc, gt_o), it
        assert torch.equal(torch.round(lq * 255) / 255, lq), it           # exact 8-bit levels
        lv = ((lq - lq_o) * 255).abs()
        frac = float((lv > 0.5).float().mean())
        worst = max(worst, frac)
        assert frac < 0.02, (it, frac)                                    # flips are isolated ...
        assert float(lv.mean()) < 0.05, (it, float(lv.mean()))            # ... and small on average
    print(f"pipeline replay: worst fraction of LQ pixels off the oracle's level = {worst:.2e}")


def test_feed_data_in_kernel_rng_runs_and_is_seeded():
    """feed_data with its own draws (no injected fields): output shape/levels/range, pool bookkeeping, and
    reproducibility from the seeds."""
    from neosr_b200.models.otf import draw_plan
    ds = dict(DEGRADATIONS, patch_size=24, batch_size=4)
    outs = []
    for rep in range(2):
        m = _otf_model(ds, 8)
        rng, pr, rd = np.random.default_rng(3), random.Random(3), np.random.default_rng(4)
        m._rng_dev = np.random.default_rng(5)
        res = []
        for it in range(5):
            gt, k1, k2, sk = case_inputs(900 + it, ds, 4, 128)
            plan = draw_plan(ds, 4, 128, 128, 4, rng, pr, rd)
            lq, g = m.run_plan(gt.to(DEV), k1.to(DEV), k2.to(DEV), sk.to(DEV), plan)
            lq, g = m._dequeue_and_enqueue(lq, g)
            assert lq.shape == (4, 3, 24, 24) and g.shape == (4, 3, 96, 96)
            assert float(lq.min()) >= 0 and float(lq.max()) <= 1
            lq, g = lq.cpu(), g.cpu()
            assert torch.equal(torch.round(lq * 255) / 255, lq)  # on the CPU: true division, as the kernel does
            res.append((lq, g))
        assert m.queue_ptr == 8
        outs.append(res)
    for (a, b), (c2, d) in zip(*outs):
        assert torch.equal(a, c2) and torch.equal(b, d)


# ---------------------------------------------------------------- apply_augment (augmentations.py:219-310)
@pytest.mark.parametrize("mode", ["bilinear", "bicubic"])
@pytest.mark.parametrize("args", [dict(scale_factor=4), dict(scale_factor=0.25), dict(size=(17, 29)), dict(size=(50, 70)), dict(scale_factor=2)])
def test_resize_antialias(mode, args):
    import torch.nn.functional as F
    img = structured_gt(9, 2, 40, 56) * 1.1 - 0.05
    ref = torch.clamp(F.interpolate(img, mode=mode, antialias=True, **args), 0, 1)
    out = ops.resize_aa(img.to(DEV), mode, **args).cpu()
    assert out.shape == ref.shape and float((out - ref).abs().max()) < 3e-6


def test_apply_augment_vs_oracle():
    """run_augment_plan on plans covering every operation and the multi branch: GT exact where no resampling is involved,
    everything within fp32 resampling noise of the oracle (== the reference, tests/test_oracle_vs_reference.py)."""
    from neosr_b200.data.augmentations import draw_augment_plan, run_augment_plan
    augs, prob = ["none", "mixup", "cutmix", "resizemix", "cutblur"], [0.5, 0.1, 0.1, 0.1, 0.5]
    seen = set()
    for seed in range(30):
        gt = structured_gt(seed, 4, 64, 64)
        lq = torch.nn.functional.avg_pool2d(gt, 4)
        plan = draw_augment_plan(4, 64, 64, 4, augs, prob, np.random.default_rng(seed), random.Random(seed), np.random.default_rng(seed + 1))
        go, lo = O.apply_augment(gt, lq, 4, plan)
        g, l = run_augment_plan(gt.to(DEV), lq.to(DEV), 4, plan)
        names = {o["op"] for o in plan["ops"]}
        seen |= names
        if "resizemix" not in names:
            assert torch.equal(g.cpu(), go), (seed, plan)
        assert float((g.cpu() - go).abs().max()) < 3e-6 and float((l.cpu() - lo).abs().max()) < 5e-6, (seed, plan)
        assert l.shape == lq.shape
    assert seen == {"mixup", "cutmix", "resizemix", "cutblur"}
    with pytest.raises(ValueError, match="batch >1"):
        draw_augment_plan(1, 64, 64, 4, augs, prob, np.random.default_rng(0))
    with pytest.raises(ValueError, match="don't match"):
        draw_augment_plan(4, 64, 64, 4, augs, prob[:3], np.random.default_rng(0))
