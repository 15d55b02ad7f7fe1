"""Pin the oracle against the LIVE reference modules (build container only).

Skipped wherever /root/reference is absent (e.g. the GPU box); the committed
fixtures in tests/golden/ carry the same check there (test_oracle_golden.py).
"""
import pytest
import torch

from oracle import losses as OL
from oracle import ref_shim
from oracle.step import make_swinir_trainer
from oracle.swinir import SwinIRConfig, swinir_forward, swinir_medium_config, swinir_param_shapes, synth_params

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.available(), reason="live reference not mounted")]

TINY = dict(img_size=16, embed_dim=36, depths=(2, 2), num_heads=(3, 3), window_size=8, mlp_ratio=2.0,
            upsampler="pixelshuffle", resi_connection="1conv", upscale=4)


def _ref_swinir(cfgkw, params):
    ref_shim.activate(4)
    from neosr.archs.swinir_arch import swinir
    kw = dict(cfgkw)
    net = swinir(drop_path_rate=0.0, **kw)
    missing = net.load_state_dict(params, strict=False)
    assert not missing.unexpected_keys
    assert all(("relative_position_index" in k or "attn_mask" in k) for k in missing.missing_keys)
    return net.train()


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("hw", [(16, 16), (24, 32)])
def test_swinir_tiny_forward_backward(hw):
    cfg = SwinIRConfig(**TINY)
    p = synth_params(swinir_param_shapes(cfg), seed=1)
    net = _ref_swinir(TINY, p)
    assert set(swinir_param_shapes(cfg)) == {k for k, _ in net.named_parameters()}
    x = torch.rand(2, 3, *hw, generator=torch.Generator().manual_seed(2))
    y_ref = net(x)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    y = swinir_forward(pr, cfg, x)
    assert _rel(y, y_ref) < 1e-5
    gt = torch.rand_like(y_ref)
    (y_ref - gt).abs().mean().backward()
    g = torch.autograd.grad((y - gt).abs().mean(), list(pr.values()))
    ref_g = dict((k, v.grad) for k, v in net.named_parameters())
    for k, gi in zip(pr, g):
        assert _rel(gi, ref_g[k]) < 2e-4, k


def test_swinir_medium_forward():
    cfg = swinir_medium_config(4)
    p = synth_params(swinir_param_shapes(cfg), seed=0)
    ref_shim.activate(4)
    net = ref_shim.build_network({"type": "swinir_medium", "drop_path_rate": 0.0})
    assert {k: tuple(v.shape) for k, v in net.named_parameters()} == swinir_param_shapes(cfg)
    net.load_state_dict(p, strict=False)
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        assert _rel(swinir_forward(p, cfg, x), net(x)) < 1e-5


def test_perceptual_and_step():
    """3 iterations of the reference's REAL optimize_parameters vs the oracle trainer."""
    cfg = SwinIRConfig(**TINY)
    p = synth_params(swinir_param_shapes(cfg), seed=4)
    vgg_p = synth_params(OL.vgg19_conv_shapes(), seed=5)
    ref_shim.activate(4)
    from neosr.losses.basic_loss import L1Loss
    net = _ref_swinir(TINY, p)
    cri_p = ref_shim.build_vgg_perceptual(vgg_p, loss_weight=0.5)
    okw = dict(lr=1e-3, betas=(0.98, 0.92, 0.987), weight_decay=0.02, schedule_free=True, warmup_steps=1600)
    model = ref_shim.make_image_model(net, cri_pix=L1Loss(1.0), cri_perceptual=cri_p, optim_kw=okw)
    tr = make_swinir_trainer(p, cfg, pixel_weight=1.0, percep_weight=0.5, vgg_params=vgg_p, optim=okw, ema=0.999)
    g = torch.Generator().manual_seed(6)
    for it in range(3):
        lq, gt = torch.rand(2, 3, 16, 16, generator=g), torch.rand(2, 3, 64, 64, generator=g)
        model.feed_data({"lq": lq, "gt": gt})
        model.optimize_parameters(it)
        tr.feed_data({"lq": lq, "gt": gt})
        tr.optimize_parameters(it)
        ref_log = model.get_current_log()
        for k, v in tr.get_current_log().items():
            assert abs(v - ref_log[k]) <= 1e-5 * max(1.0, abs(ref_log[k])), (it, k, v, ref_log[k])
    ref_params = dict(net.named_parameters())
    ema_params = dict(model.net_g_ema.module.named_parameters())
    for i, k in enumerate(tr.names):
        assert _rel(tr.params[k].detach(), ref_params[k].detach()) < 1e-4, k
        assert _rel(tr.ema.avg[i], ema_params[k].detach()) < 1e-4, k
    sd = model.optimizer_g.state_dict()
    assert set(sd["state"][0]) == {"exp_avg", "exp_avg_sq", "exp_avg_diff", "z", "neg_pre_grad"}
    assert abs(sd["param_groups"][0]["weight_sum"] - tr.opt.weight_sum) < 1e-12


def test_compact_forward_backward():
    from oracle.compact import compact_forward, compact_param_shapes
    ref_shim.activate(4)
    net = ref_shim.build_network({"type": "compact", "upscale": 2, "num_conv": 4, "num_feat": 32})
    shapes = compact_param_shapes(num_feat=32, num_conv=4, upscale=2)
    assert {k: tuple(v.shape) for k, v in net.named_parameters()} == shapes
    p = synth_params(shapes, seed=7)
    net.load_state_dict(p)
    x = torch.rand(2, 3, 16, 24, generator=torch.Generator().manual_seed(8))
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    y, y_ref = compact_forward(pr, x, num_conv=4, upscale=2), net(x)
    assert _rel(y, y_ref) < 1e-5
    (y_ref ** 2).mean().backward()
    g = torch.autograd.grad((y ** 2).mean(), list(pr.values()))
    for (k, v), gi in zip(net.named_parameters(), g):
        assert _rel(gi, v.grad) < 1e-4, k


@pytest.mark.parametrize("scale", [4, 2])
def test_esrgan_forward_backward(scale):
    from oracle.esrgan import esrgan_forward, esrgan_param_shapes
    ref_shim.activate(4)
    net = ref_shim.build_network({"type": "esrgan", "scale": scale, "num_block": 2, "num_feat": 32, "num_grow_ch": 16})
    shapes = esrgan_param_shapes(scale=scale, num_feat=32, num_block=2, num_grow_ch=16)
    assert {k: tuple(v.shape) for k, v in net.named_parameters()} == shapes
    p = synth_params(shapes, seed=11)
    net.load_state_dict(p)
    x = torch.rand(2, 3, 16, 24, generator=torch.Generator().manual_seed(12))
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    y, y_ref = esrgan_forward(pr, x, scale=scale, num_block=2), net(x)
    assert _rel(y, y_ref) < 1e-5
    (y_ref ** 2).mean().backward()
    g = torch.autograd.grad((y ** 2).mean(), list(pr.values()))
    for (k, v), gi in zip(net.named_parameters(), g):
        assert _rel(gi, v.grad) < 1e-4, k


def test_unet_sn_forward_backward_and_buffers():
    """Three training forwards (as one GAN step makes) + an eval forward: outputs, gradients w.r.t.
    weight_orig / input, and the evolution of the spectral-norm u/v buffers match the reference module."""
    from oracle.unet import synth_unet, unet_forward, unet_param_shapes
    ref_shim.activate(4)
    net = ref_shim.build_network({"type": "unet", "num_feat": 16})
    ps, bs = unet_param_shapes(num_feat=16)
    assert {k: tuple(v.shape) for k, v in net.named_parameters()} == ps
    assert {k: tuple(v.shape) for k, v in net.named_buffers()} == bs
    p, b = synth_unet(num_feat=16, seed=21)
    net.load_state_dict({**p, **b})
    net.train()
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    bo = {k: v.clone() for k, v in b.items()}
    gen = torch.Generator().manual_seed(22)
    for it in range(3):
        x = torch.rand(2, 3, 32, 40, generator=gen)
        xo, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        y, y_ref = unet_forward(pr, bo, xo, training=True), net(xr)
        assert _rel(y, y_ref) < 1e-5, it
        for k, v in net.named_buffers():
            assert _rel(bo[k], v) < 1e-5, (it, k)
        net.zero_grad()
        (y_ref ** 2).mean().backward()
        g = torch.autograd.grad((y ** 2).mean(), [*pr.values(), xo])
        for (k, v), gi in zip(net.named_parameters(), g):
            assert _rel(gi, v.grad) < 1e-4, (it, k)
        assert _rel(g[-1], xr.grad) < 1e-4
    net.eval()
    x = torch.rand(1, 3, 16, 16, generator=gen)
    with torch.no_grad():
        assert _rel(unet_forward(pr, bo, x, training=False), net(x)) < 1e-5


def test_gan_step_with_unet_discriminator():
    """C2's step shape: 3 iterations of the reference's REAL optimize_parameters with net_d = unet and
    gan_loss(bce) vs the oracle trainer — logs, both networks' parameters, spectral-norm buffers."""
    from oracle.unet import synth_unet
    cfg = SwinIRConfig(**TINY)
    p = synth_params(swinir_param_shapes(cfg), seed=4)
    dp, db = synth_unet(num_feat=16, seed=9)
    ref_shim.activate(4)
    from neosr.losses.basic_loss import L1Loss
    from neosr.losses.gan_loss import gan_loss
    net = _ref_swinir(TINY, p)
    net_d = ref_shim.build_network({"type": "unet", "num_feat": 16})
    net_d.load_state_dict({**dp, **db})
    okw = dict(lr=1e-3, betas=(0.98, 0.92, 0.987), weight_decay=0.02, schedule_free=True, warmup_steps=1600)
    model = ref_shim.make_image_model(net, cri_pix=L1Loss(1.0), optim_kw=okw, net_d=net_d,
                                      cri_gan=gan_loss("bce", loss_weight=0.3))
    tr = make_swinir_trainer(p, cfg, pixel_weight=1.0, optim=okw, ema=0.999, disc=(dp, db), gan_weight=0.3)
    g = torch.Generator().manual_seed(6)
    for it in range(3):
        lq, gt = torch.rand(2, 3, 16, 16, generator=g), torch.rand(2, 3, 64, 64, generator=g)
        model.feed_data({"lq": lq, "gt": gt})
        model.optimize_parameters(it)
        tr.feed_data({"lq": lq, "gt": gt})
        tr.optimize_parameters(it)
        ref_log = model.get_current_log()
        olog = tr.get_current_log()
        assert set(olog) == set(ref_log), (sorted(olog), sorted(ref_log))
        for k, v in olog.items():
            assert abs(v - ref_log[k]) <= 1e-5 * max(1.0, abs(ref_log[k])), (it, k, v, ref_log[k])
    ref_params = dict(net.named_parameters())
    for k in tr.names:
        assert _rel(tr.params[k].detach(), ref_params[k].detach()) < 1e-4, k
    for k, v in net_d.named_parameters():
        assert _rel(tr.d_params[k].detach(), v.detach()) < 1e-4, k
    for k, v in net_d.named_buffers():
        assert _rel(tr.d_buffers[k], v) < 1e-5, k


@pytest.mark.parametrize("use_ea,ks", [(True, 17), (False, 13)])
def test_realplksr_forward_backward(use_ea, ks):
    from oracle.realplksr import realplksr_forward, realplksr_param_shapes
    ref_shim.activate(4)
    kw = dict(dim=32, n_blocks=2, upscaling_factor=4, kernel_size=ks, use_ea=use_ea)
    net = ref_shim.build_network({"type": "realplksr", **kw}).train()
    shapes = realplksr_param_shapes(**kw)
    assert {k: tuple(v.shape) for k, v in net.named_parameters()} == shapes
    p = synth_params(shapes, seed=13)
    net.load_state_dict(p)
    x = torch.rand(2, 3, 20, 24, generator=torch.Generator().manual_seed(14))
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    y, y_ref = realplksr_forward(pr, x, n_blocks=2, kernel_size=ks, use_ea=use_ea), net(x)
    assert _rel(y, y_ref) < 1e-5
    (y_ref ** 2).mean().backward()
    g = torch.autograd.grad((y ** 2).mean(), list(pr.values()))
    for (k, v), gi in zip(net.named_parameters(), g):
        assert _rel(gi, v.grad) < 1e-4, k


# ---------------------------------------------------------------- OTF degradation pipeline
def test_otf_stage_functions_vs_reference():
    """oracle.otf.{filter2d, jpeg, gaussian_noise, poisson_noise} against the reference's own functions
    (diffjpeg.py:531-584, degradations.py:569-605,738-786) on the same inputs and the same random draws."""
    from oracle import otf as O
    from oracle.ref_otf import structured_gt
    ref_shim.activate(4)
    import neosr.utils.diffjpeg as dj
    dj.device = torch.device("cpu")
    from neosr.data import degradations as D
    g = torch.Generator().manual_seed(0)
    img = structured_gt(3, 3, 50, 70)
    k = torch.rand(3, 21, 21, generator=g)
    k /= k.sum((1, 2), keepdim=True)
    assert torch.equal(O.filter2d(img, k), dj.filter2D(img, k))
    k1 = torch.rand(1, 7, 7, generator=g)
    assert torch.equal(O.filter2d(img, k1), dj.filter2D(img, k1))
    with pytest.raises(ValueError, match="Wrong kernel size"):
        dj.filter2D(img, torch.rand(1, 8, 8))
    with pytest.raises(ValueError, match="Wrong kernel size"):
        O.filter2d(img, torch.rand(1, 8, 8))
    q = torch.tensor([35.0, 50.0, 93.0])
    ref = dj.DiffJPEG(differentiable=False)(img.clone(), quality=q.clone())
    assert float((ref - O.jpeg(img, q)).abs().max()) < 1e-6
    z, zg = torch.randn(3, 3, 50, 70, generator=g), torch.randn(50, 70, generator=g)
    sigma, gray = torch.tensor([3.0, 10.0, 25.0]), torch.tensor([0.0, 1.0, 0.0])
    calls, orig = [zg, z], torch.randn
    torch.randn = lambda *a, **kw: calls.pop(0)
    try:
        noise = D.generate_gaussian_noise_pt(img, sigma, gray)
    finally:
        torch.randn = orig
    assert torch.equal(torch.clamp(img + noise, 0, 1), O.gaussian_noise(img, sigma, gray, z, zg))
    rec, origp = [], torch.poisson

    def fakep(lam, generator=None):
        rec.append(origp(lam, generator=g))
        return rec[-1]

    x = img * 0.7 + 0.1 * torch.rand(3, 3, 50, 70, generator=g)
    scale = torch.tensor([0.5, 1.0, 2.0])
    torch.poisson = fakep
    try:
        noise = D.generate_poisson_noise_pt(x, scale, gray)
    finally:
        torch.poisson = origp
    out, _, _ = O.poisson_noise(x, scale, gray, counts_color=rec[1], counts_gray=rec[0])
    assert torch.equal(torch.clamp(x + noise, 0, 1), out)


def test_otf_feed_data_record_and_replay():
    """The reference's REAL otf.feed_data on CPU (oracle/ref_otf.py), recorded, then replayed through
    oracle.otf.degrade + Pool: bit-identical LQ/GT over iterations that fill and then cycle the pool."""
    import random

    import numpy as np

    from neosr_b200.data.degradations import synth_kernels
    from oracle import otf as O
    from oracle import ref_otf as R
    ds = dict(R.DEGRADATIONS, patch_size=16, batch_size=2)
    model, pool = None, O.Pool(4)
    for it in range(5):
        seed = 500 + it
        gt = R.structured_gt(seed, 2, 96, 96)
        rng, pr = np.random.default_rng(seed), random.Random(seed)
        ks = [synth_kernels(ds, rng, pr) for _ in range(2)]
        k1, k2, sk = [torch.from_numpy(np.stack([k[i] for k in ks])) for i in range(3)]
        lq_r, gt_r, plan, fields, perm, model = R.run_reference(gt, k1, k2, sk, ds, 4, seed, model=model, queue_size=4)
        lq, gtc, _, _ = O.degrade(gt, k1, k2, sk, plan, 4, fields)
        lq, gtc = pool.step(lq, gtc, perm)
        assert torch.equal(lq, lq_r) and torch.equal(gtc, gt_r), it
        assert torch.equal(torch.round(lq * 255) / 255, lq)


def test_mssim_and_consistency_vs_reference_modules():
    """oracle.losses.{msssim_loss, consistency_loss} vs the live reference modules: value and input gradient, including
    the case where the data-dependent cosine branch (consistency_loss.py:186-190) is active."""
    from oracle.make_golden_otf import loss_inputs
    ref_shim.activate(4)
    from neosr.losses.consistency_loss import consistency_loss
    from neosr.losses.ssim_loss import mssim_loss
    for tag, (x, gt) in loss_inputs().items():
        x = x.clone().requires_grad_(True)
        cases = [(mssim_loss(loss_weight=0.7), lambda a, b: OL.msssim_loss(a, b, 0.7)),
                 (consistency_loss(), lambda a, b: OL.consistency_loss(a, b)),
                 (consistency_loss(blur=False, saturation=1.2, brightness=0.9, loss_weight=0.5),
                  lambda a, b: OL.consistency_loss(a, b, 0.5, blur=False, saturation=1.2, brightness=0.9))]
        for mod, fn in cases:
            r, o = mod(x, gt), fn(x, gt)
            gr, = torch.autograd.grad(r, x)
            go, = torch.autograd.grad(o, x)
            assert float(r.detach()) == float(o.detach()), tag
            assert _rel(go, gr) < 1e-6, tag


def test_hat_oracle_and_module_surface_vs_reference():
    """oracle.hat.hat_forward (forward + every parameter gradient) against the live reference `hat`; and the product
    module's state_dict: same keys, order, shapes and index buffers as the reference (checkpoints / optimizer states
    interchange)."""
    from neosr_b200.archs.hat_arch import hat as our_hat
    from oracle.hat import HATConfig, hat_forward, hat_param_shapes
    ref_shim.activate(4)
    from neosr.archs.hat_arch import hat
    kw = dict(img_size=64, embed_dim=36, depths=(2, 2), num_heads=(3, 3), window_size=16, compress_ratio=3, squeeze_factor=6,
              conv_scale=0.01, overlap_ratio=0.5, mlp_ratio=2, upscale=4)
    net = hat(drop_path_rate=0.0, upsampler="pixelshuffle", resi_connection="1conv", **kw).train()
    cfg = HATConfig(**kw)
    shapes = hat_param_shapes(cfg)
    assert shapes == {k: tuple(v.shape) for k, v in net.named_parameters()}
    p = synth_params(shapes, seed=3)
    net.load_state_dict(p, strict=False)
    x = torch.rand(2, 3, 32, 48, generator=torch.Generator().manual_seed(1))
    y_ref = net(x)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    y = hat_forward(pr, cfg, x)
    assert _rel(y.detach(), y_ref.detach()) < 1e-5
    gt = torch.rand_like(y_ref)
    ((y_ref - gt) ** 2).mean().backward()
    g = torch.autograd.grad(((y - gt) ** 2).mean(), list(pr.values()))
    rg = {k: v.grad for k, v in net.named_parameters()}
    for k, gi in zip(pr, g):
        assert _rel(gi, rg[k]) < 2e-4, k
    ours = our_hat(drop_path_rate=0.0, upsampler="pixelshuffle", resi_connection="1conv", **kw)
    assert [(k, tuple(v.shape)) for k, v in ours.state_dict().items()] == [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert torch.equal(ours.relative_position_index_SA, net.relative_position_index_SA)
    assert torch.equal(ours.relative_position_index_OCA, net.relative_position_index_OCA)
    assert int(net.relative_position_index_OCA.min()) < 0  # the reference's OCA index has negative (wrapping) entries


def test_apply_augment_plan_and_oracle_vs_reference():
    """neosr_b200.data.augmentations.draw_augment_plan draws what the reference's apply_augment draws (same python
    `random` / numpy streams; torch.randperm results injected), and oracle.otf.apply_augment then reproduces the
    reference's output bit for bit, over seeds that cover every operation and the multi-augmentation branch."""
    import random

    import numpy as np

    from neosr_b200.data.augmentations import draw_augment_plan
    from neosr_b200.data.synthetic import structured_gt
    from oracle import otf as O
    ref_shim.activate(4)
    import neosr.data.augmentations as A
    augs, prob = ["none", "mixup", "cutmix", "resizemix", "cutblur"], [0.5, 0.1, 0.1, 0.1, 0.5]
    seen = set()
    for seed in range(40):
        gt = structured_gt(seed, 4, 64, 64)
        lq = torch.nn.functional.avg_pool2d(gt, 4)
        A.rng, A.random = np.random.default_rng(seed), random.Random(seed)
        perms, o_randperm, g = [], torch.randperm, torch.Generator().manual_seed(seed)

        def rp(n, **k):
            perms.append(o_randperm(n, generator=g))
            return perms[-1]

        torch.randperm = rp
        try:
            gr, lr = A.apply_augment(gt.clone(), lq.clone(), scale=4, augs=augs, prob=prob)
        finally:
            torch.randperm = o_randperm

        class Replay:
            i = 0

            def permutation(self, n):
                self.i += 1
                return perms[self.i - 1].numpy()

        plan = draw_augment_plan(4, 64, 64, 4, augs, prob, np.random.default_rng(seed), random.Random(seed), Replay())
        go, lo = O.apply_augment(gt, lq, 4, plan)
        assert torch.equal(go, gr) and torch.equal(lo, lr), (seed, plan)
        seen |= {o["op"] for o in plan["ops"]}
    assert seen == {"mixup", "cutmix", "resizemix", "cutblur"}


def test_swinir_3conv_nearest_conv_oracle_and_keys_vs_reference():
    """swinir_large's variants (3conv, nearest+conv): oracle forward == reference, product state_dict == reference's."""
    from neosr_b200.archs.swinir_arch import swinir as our_swinir
    ref_shim.activate(4)
    from neosr.archs.swinir_arch import swinir
    kw = dict(img_size=16, embed_dim=48, depths=(2, 2), num_heads=(4, 4), window_size=8, mlp_ratio=2.0, upsampler="nearest+conv",
              resi_connection="3conv", upscale=4)
    ref, ours = swinir(drop_path_rate=0.0, **kw).train(), our_swinir(drop_path_rate=0.0, **kw)
    assert [(k, tuple(v.shape)) for k, v in ref.state_dict().items()] == [(k, tuple(v.shape)) for k, v in ours.state_dict().items()]
    cfg = SwinIRConfig(**kw)
    p = synth_params(swinir_param_shapes(cfg), seed=9)
    ref.load_state_dict(p, strict=False)
    x = torch.rand(2, 3, 16, 24, generator=torch.Generator().manual_seed(0))
    assert _rel(swinir_forward(p, cfg, x), ref(x).detach()) < 1e-5


def test_validation_host_helpers_vs_reference():
    """tensor2img / calculate_psnr of the standalone fallback (neosr_b200/models/_validation.py) against the reference's
    own (`utils/img_util.py:60-129`, `metrics/calculate.py:16-66`), incl. the clamp, grayscale and crop-border paths."""
    import importlib
    import sys

    import numpy as np

    ref_shim.activate(4)
    from neosr.metrics.calculate import calculate_psnr as ref_psnr
    from neosr.utils import tensor2img as ref_t2i

    # the fallbacks, not the delegating wrappers: import the module without `neosr` shadowing its own functions
    V = importlib.import_module("neosr_b200.models._validation")
    g = torch.Generator().manual_seed(11)
    a = torch.rand(1, 3, 40, 56, generator=g) * 1.2 - 0.1  # values outside [0, 1] exercise the clamp
    b = (a + 0.05 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    ia, ib = V.tensor2img(a), V.tensor2img(b)
    ra, rb = ref_t2i(a), ref_t2i(b)
    assert ia.dtype == np.uint8 and ia.shape == ra.shape and np.array_equal(ia, ra) and np.array_equal(ib, rb)
    gray = torch.rand(1, 1, 24, 24, generator=g)
    assert np.array_equal(V.tensor2img(gray), ref_t2i(gray))
    # (test_y_channel=True is not compared live: the reference's to_y_channel raises under numpy 2.x, whose dtype objects
    #  are not members of the {np.float32, np.float16} set it tests - utils/color_util.py:177-186)
    for kw in (dict(crop_border=4, test_y_channel=False), dict(crop_border=0, test_y_channel=False),
               dict(crop_border=7, test_y_channel=False)):
        ours, ref = V.calculate_psnr(ia, ib, **kw), float(ref_psnr(ra, rb, **kw))
        assert abs(ours - ref) <= 1e-4 * abs(ref), (kw, ours, ref)
    y = V.calculate_psnr(ia, ib, crop_border=4, test_y_channel=True)
    assert np.isfinite(y) and y > V.calculate_psnr(ia, ib, crop_border=4) - 10.0
    assert V.calculate_psnr(ia, ia) == float("inf")
    assert "neosr_b200.models._validation" in sys.modules


def test_eco_and_fsam_step_variants_vs_reference():
    """Opt-in step variants: the oracle's ECO centroid step and F-SAM double closure against the
    reference's REAL closure / optimize_parameters (image.py:393-425, 627-662; optimizers/fsam.py)."""
    ref_shim.activate(4)
    from neosr.losses.basic_loss import L1Loss
    from oracle.make_golden import TINY
    from oracle.step import make_swinir_trainer
    cfg = SwinIRConfig(**TINY)
    okw = dict(lr=1e-3, betas=(0.98, 0.92, 0.987), weight_decay=0.02, schedule_free=True, warmup_steps=0)
    for variant in ("eco", "fsam"):
        p = synth_params(swinir_param_shapes(cfg), seed=21)
        net = _ref_swinir(TINY, p)
        eco = dict(iters=8, init=2, schedule="sigmoid", pretrain=None) if variant == "eco" else None
        model = ref_shim.make_image_model(net, cri_pix=L1Loss(1.0), optim_kw=okw, eco=eco, sam_init=0 if variant == "fsam" else None)
        tr = make_swinir_trainer(p, cfg, pixel_weight=1.0, optim=okw, ema=0.999, eco=eco,
                                 sam=dict(init=0) if variant == "fsam" else None, scale=4)
        g = torch.Generator().manual_seed(22)
        for it in range(1, 5):
            lq, gt = torch.rand(2, 3, 16, 16, generator=g), torch.rand(2, 3, 64, 64, generator=g)
            model.feed_data({"lq": lq, "gt": gt})
            model.optimize_parameters(it)
            tr.feed_data({"lq": lq, "gt": gt})
            tr.optimize_parameters(it)
            ref_log = model.get_current_log()
            for k, v in tr.get_current_log().items():
                assert abs(v - ref_log[k]) <= 1e-5 * max(1.0, abs(ref_log[k])), (variant, it, k, v, ref_log[k])
        ref_params = dict(net.named_parameters())
        for k in tr.names:
            assert _rel(tr.params[k].detach(), ref_params[k].detach()) < 1e-4, (variant, k)
