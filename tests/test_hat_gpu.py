"""HAT on the B200 kernels (neosr_b200/archs/hat_arch.py, csrc/hat_ops.cu) against the CPU oracle (oracle/hat.py,
pinned to the reference module): the (cross-)window attention kernels alone, the whole network forward + every
parameter gradient on both contraction engines, and C4-shaped training steps."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from neosr_b200 import ops  # noqa: E402
from oracle.hat import HATConfig, hat_forward, hat_param_shapes, rpi_oca, rpi_sa  # noqa: E402
from oracle.swinir import calculate_mask, synth_params, window_partition, window_reverse  # noqa: E402

TINY = dict(img_size=64, embed_dim=36, depths=(2, 2), num_heads=(3, 3), window_size=16, compress_ratio=3, squeeze_factor=6,
            conv_scale=0.01, overlap_ratio=0.5, mlp_ratio=2, upscale=4)


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


class _engine:
    """Select the nsr_xwin_attn_* engine (1 tensor cores, 0 exact fp32) for a block."""

    def __init__(self, tc):
        self.tc = tc

    def __enter__(self):
        self.prev, ops.XWIN_TENSOR_CORES = ops.XWIN_TENSOR_CORES, int(self.tc)

    def __exit__(self, *a):
        ops.XWIN_TENSOR_CORES = self.prev


def _self_attn_ref(qkv, table, heads, ws, shift):
    """HAB attention core from a [B,H,W,3C] qkv tensor: roll, partition, softmax(QK^T*scale + bias + mask) V, reverse."""
    B, H, W, c3 = qkv.shape
    c = c3 // 3
    x = torch.roll(qkv, shifts=(-shift, -shift), dims=(1, 2)) if shift else qkv
    xw = window_partition(x, ws).view(-1, ws * ws, 3, heads, c // heads).permute(2, 0, 3, 1, 4)
    q, k, v = xw[0] * (c // heads) ** -0.5, xw[1], xw[2]
    attn = q @ k.transpose(-2, -1)
    bias = table[rpi_sa(ws).view(-1)].view(ws * ws, ws * ws, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if shift:
        mask = calculate_mask(H, W, ws, shift).to(qkv.dtype)
        nw = mask.shape[0]
        attn = (attn.view(-1, nw, heads, ws * ws, ws * ws) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, ws * ws, ws * ws)
    o = (torch.softmax(attn, -1) @ v).transpose(1, 2).reshape(-1, ws, ws, c)
    o = window_reverse(o, ws, H, W)
    return torch.roll(o, shifts=(shift, shift), dims=(1, 2)) if shift else o


def _oca_ref(qkv, table, heads, ws, ows):
    B, H, W, c3 = qkv.shape
    c = c3 // 3
    q = qkv[..., :c]
    kv = qkv[..., c:].permute(0, 3, 1, 2)  # b, 2c, h, w  (k | v)
    qw = window_partition(q, ws).view(-1, ws * ws, c)
    kvw = F.unfold(kv, kernel_size=(ows, ows), stride=ws, padding=(ows - ws) // 2)
    nw = kvw.shape[-1]
    kvw = kvw.view(B, 2, c, ows, ows, nw).permute(1, 0, 5, 3, 4, 2).reshape(2, B * nw, ows * ows, c)
    d = c // heads
    qh = qw.reshape(-1, ws * ws, heads, d).permute(0, 2, 1, 3) * d ** -0.5
    kh = kvw[0].reshape(-1, ows * ows, heads, d).permute(0, 2, 1, 3)
    vh = kvw[1].reshape(-1, ows * ows, heads, d).permute(0, 2, 1, 3)
    bias = table[rpi_oca(ws, (ows - ws) / ws).view(-1)].view(ws * ws, ows * ows, -1).permute(2, 0, 1)
    attn = torch.softmax(qh @ kh.transpose(-2, -1) + bias.unsqueeze(0), -1)
    o = (attn @ vh).transpose(1, 2).reshape(-1, ws, ws, c)
    return window_reverse(o, ws, H, W)


@pytest.mark.parametrize("case", [dict(B=2, H=32, W=48, c=36, heads=3, ws=16, shift=0), dict(B=2, H=32, W=48, c=36, heads=3, ws=16, shift=8),
                                  dict(B=1, H=32, W=32, c=180, heads=6, ws=16, shift=8), dict(B=1, H=16, W=24, c=24, heads=2, ws=8, shift=4)])
@pytest.mark.parametrize("tc", [0, 3])
def test_xwin_attn_self(case, tc):
    """tc = 0: exact-fp32 CUDA-core kernels (tight bounds); tc = 3: mma.sync 3xBF16 forward AND backward (1e-4 class)."""
    with _engine(tc):
        _xwin_self(case, 1.0 if tc == 0 else 10.0)


def _xwin_self(case, loosen):
    B, H, W, c, heads, ws, shift = (case[k] for k in ("B", "H", "W", "c", "heads", "ws", "shift"))
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B, H, W, 3 * c, generator=g).requires_grad_(True)
    table = (0.3 * torch.randn((2 * ws - 1) ** 2, heads, generator=g)).requires_grad_(True)
    dout = torch.randn(B, H, W, c, generator=g)
    ref = _self_attn_ref(qkv, table, heads, ws, shift)
    gq, gt = torch.autograd.grad((ref * dout).sum(), [qkv, table])
    scale = (c // heads) ** -0.5
    out, lse = ops.xwin_attn_fwd(qkv.detach().cuda(), table.detach().cuda(), heads, ws, ws, shift, scale)
    assert rel(out, ref) < 2e-5 * loosen
    dtab = torch.empty_like(table.detach()).cuda()
    dqkv = ops.xwin_attn_bwd(qkv.detach().cuda(), table.detach().cuda(), out, dout.cuda(), lse, dtab, heads, ws, ws, shift, scale)
    assert rel(dqkv, gq) < 5e-5 * loosen and rel(dtab, gt) < 5e-5 * loosen
    dtab2 = torch.empty_like(dtab)
    dqkv2 = ops.xwin_attn_bwd(qkv.detach().cuda(), table.detach().cuda(), out, dout.cuda(), lse, dtab2, heads, ws, ws, shift, scale)
    assert torch.equal(dqkv, dqkv2) and torch.equal(dtab, dtab2)  # deterministic


@pytest.mark.parametrize("case", [dict(B=2, H=32, W=48, c=36, heads=3, ws=16, ows=24), dict(B=1, H=32, W=32, c=180, heads=6, ws=16, ows=24),
                                  dict(B=2, H=16, W=16, c=24, heads=2, ws=16, ows=24), dict(B=1, H=16, W=24, c=24, heads=2, ws=8, ows=12)])
@pytest.mark.parametrize("tc", [0, 3])
def test_xwin_attn_overlapping(case, tc):
    with _engine(tc):
        _xwin_oca(case, 1.0 if tc == 0 else 10.0)


def _xwin_oca(case, loosen):
    B, H, W, c, heads, ws, ows = (case[k] for k in ("B", "H", "W", "c", "heads", "ws", "ows"))
    g = torch.Generator().manual_seed(4)
    qkv = torch.randn(B, H, W, 3 * c, generator=g).requires_grad_(True)
    table = (0.3 * torch.randn((ws + ows - 1) ** 2, heads, generator=g)).requires_grad_(True)
    dout = torch.randn(B, H, W, c, generator=g)
    ref = _oca_ref(qkv, table, heads, ws, ows)
    gq, gt = torch.autograd.grad((ref * dout).sum(), [qkv, table])
    scale = (c // heads) ** -0.5
    out, lse = ops.xwin_attn_fwd(qkv.detach().cuda(), table.detach().cuda(), heads, ws, ows, 0, scale)
    assert rel(out, ref) < 2e-5 * loosen
    dtab = torch.empty_like(table.detach()).cuda()
    dqkv = ops.xwin_attn_bwd(qkv.detach().cuda(), table.detach().cuda(), out, dout.cuda(), lse, dtab, heads, ws, ows, 0, scale)
    assert rel(dqkv, gq) < 5e-5 * loosen and rel(dtab, gt) < 5e-5 * loosen


def test_xwin_attn_errors():
    qkv = torch.randn(1, 32, 32, 3 * 36, device="cuda")
    tab = torch.zeros(31 * 31, 3, device="cuda")
    with pytest.raises(RuntimeError, match="multiple of window"):
        ops.xwin_attn_fwd(torch.randn(1, 30, 32, 108, device="cuda"), tab, 3, 16, 16, 0, 1.0)
    with pytest.raises(RuntimeError, match="shift_size must in 0-window_size"):
        ops.xwin_attn_fwd(qkv, tab, 3, 16, 16, 16, 1.0)
    with pytest.raises(RuntimeError, match="no shift"):
        ops.xwin_attn_fwd(qkv, torch.zeros(39 * 39, 3, device="cuda"), 3, 16, 24, 8, 1.0)


def _tiny_net(seed=31):
    from neosr_b200.archs.hat_arch import hat
    cfg = HATConfig(**TINY)
    p = synth_params(hat_param_shapes(cfg), seed=seed)
    net = hat(drop_path_rate=0.0, upsampler="pixelshuffle", resi_connection="1conv", **TINY).cuda().train()
    missing = net.load_state_dict(p, strict=False)
    assert not missing.unexpected_keys and set(missing.missing_keys) == {"relative_position_index_SA", "relative_position_index_OCA"}
    return net, cfg, p


@pytest.mark.parametrize("hw", [(32, 32), (16, 48)])
def test_hat_tiny_forward_backward(hw):
    net, cfg, p = _tiny_net()
    g = torch.Generator().manual_seed(32)
    x = torch.rand(2, 3, *hw, generator=g)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    y_ref = hat_forward(pr, cfg, x)
    gt = torch.rand(y_ref.shape, generator=g)
    grads = torch.autograd.grad(((y_ref - gt) ** 2).mean(), list(pr.values()))
    # Smooth loss, so the north-star bound (1e-3) holds on both engines.  The exact-fp32 engine is held to 5e-4, not tighter:
    # the network has kinks (ReLU in the channel attention, LeakyReLU after conv_before_upsample), and in the hw1 case a unit
    # sits close enough to one that ANY re-ordering of an upstream fp32 sum (LayerNorm rows, the pooling mean, the bias-gradient
    # reduction - three unrelated kernel changes produced the same picture) flips it against the torch reference: single
    # tensors then move from < 2e-4 to 2.3e-4 ... 3.3e-4 while the forward output stays within 1e-4.
    for engine, tol in (("simt", 5e-4), ("auto", 1e-3)):
        ops.DEFAULT_ENGINE = engine
        try:
            net.zero_grad()
            y = net(x.cuda())
            assert rel(y, y_ref) < 1e-4
            ((y - gt.cuda()) ** 2).mean().backward()
        finally:
            ops.DEFAULT_ENGINE = "auto"
        ref_g = dict(zip(pr, grads))
        for k, v in net.named_parameters():
            # bias tables: each entry sums dS over every window and head position with heavy cancellation (|g| ~ 1e-6), so
            # upstream differences show up amplified on either engine (measured: 1.5e-3 on the OCAB table for the split-bf16
            # engine; 5.1e-4 on a HAB table for the exact engine after the flip described above)
            t = 3 * tol if k.endswith("relative_position_bias_table") else tol
            assert rel(v.grad, ref_g[k]) < t, (engine, k, rel(v.grad, ref_g[k]))


def test_hat_registry_names_and_eval_forward():
    from neosr_b200.registry import ARCH_REGISTRY
    for n in ("hat_s", "hat_m", "hat_l"):
        assert n in ARCH_REGISTRY
    net, cfg, p = _tiny_net()
    net.eval()
    x = torch.rand(1, 3, 16, 16, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        y = net(x.cuda())
    assert rel(y, hat_forward(p, cfg, x)) < 1e-4 and y.shape == (1, 3, 64, 64)
    with pytest.raises(ValueError, match="multiple of window_size"):
        net(torch.rand(1, 3, 24, 16).cuda())


def test_hat_drop_path_matches_oracle_with_same_masks():
    """DropPath (arch_util.py:118-131) on the attention and MLP branches: same per-sample factors on both sides."""
    from neosr_b200.archs.hat_arch import hat
    cfg = HATConfig(**TINY)
    p = synth_params(hat_param_shapes(cfg), seed=5)
    net = hat(drop_path_rate=0.5, upsampler="pixelshuffle", resi_connection="1conv", **TINY).cuda().train()
    net.load_state_dict(p, strict=False)
    x = torch.rand(4, 3, 16, 16, generator=torch.Generator().manual_seed(2))
    torch.manual_seed(123)
    y, S = net.engine_forward(x.cuda(), save=True)
    drops = []
    for blocks_saved, _, _ in S["layers"]:
        for rec in blocks_saved:
            ds = rec[-1]
            drops.append(None if ds is None else tuple(t.cpu() for t in ds))
    assert any(d is not None and float(d[0].min()) == 0.0 for d in drops)
    y_ref = hat_forward(p, cfg, x, drop_scales=[d if d is not None else (torch.ones(4), torch.ones(4)) for d in drops])
    assert rel(y, y_ref) < 1e-4


def test_c4_shaped_training_steps_vs_oracle():
    """hat (tiny) x4, template loss stack without the networks that need downloads: L1 + MS-SSIM + consistency,
    adan_sf + EMA — 3 iterations of the `image` model against the oracle trainer."""
    from neosr_b200.models import build_model
    from neosr_b200.registry import ARCH_REGISTRY
    from neosr_b200.archs.hat_arch import hat
    from oracle.ref_otf import structured_gt
    from oracle.step import OracleTrainer
    if "hat" not in ARCH_REGISTRY:
        ARCH_REGISTRY.register(hat)
    cfg = HATConfig(**TINY)
    p = synth_params(hat_param_shapes(cfg), seed=41)
    okw = dict(lr=1e-3, betas=(0.98, 0.92, 0.987), weight_decay=0.02, schedule_free=True, warmup_steps=1600)
    opt = {"model_type": "image", "scale": 4, "is_train": True, "dist": False, "rank": 0, "world_size": 1,
           "network_g": {"type": "hat", "drop_path_rate": 0.0, "upsampler": "pixelshuffle", "resi_connection": "1conv", **TINY},
           "datasets": {"train": {"patch_size": 16}},
           "train": {"ema": 0.999, "optim_g": {"type": "adan_sf", **okw}, "pixel_opt": {"type": "L1Loss", "loss_weight": 1.0},
                     "mssim_opt": {"type": "mssim_loss", "loss_weight": 1.0},
                     "consistency_opt": {"type": "consistency_loss", "loss_weight": 1.0}}, "path": {}}
    model = build_model(opt)
    model.net_g.load_state_dict(p, strict=False)
    tr = OracleTrainer(p, lambda q, x: hat_forward(q, cfg, x), pixel_weight=1.0, mssim_weight=1.0, consistency_weight=1.0,
                       optim=okw, ema=0.999)
    for it in range(3):
        gt = structured_gt(50 + it, 2, 64, 64)
        lq = F.avg_pool2d(gt, 4)
        model.feed_data({"lq": lq, "gt": gt})
        model.optimize_parameters(it)
        tr.feed_data({"lq": lq, "gt": gt})
        tr.optimize_parameters(it)
        log = model.get_current_log()
        for k, v in tr.get_current_log().items():
            assert abs(log[k] - v) <= 1e-3 * max(1e-3, abs(v)), (it, k, log[k], v)
    for k, v in model.net_g.named_parameters():
        assert rel(v, tr.params[k]) < 2e-3, (k, rel(v, tr.params[k]))
