"""Per-kernel parity on the GPU: every C-ABI kernel vs the same op in plain fp32 PyTorch
(TF32 off) or vs the oracle, on seeded inputs.  Tolerances: 1e-4 relative (max-abs / max-abs)
for fp32 contractions, bit-exact for index maps."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


ENGINES = ["simt", "tcgen05", "auto"]  # auto also routes 3-channel convs to conv_small.cu


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("B,H,W,cin,cout,k", [
    (2, 16, 24, 36, 36, 3), (1, 8, 8, 3, 180, 3), (2, 16, 16, 64, 3, 3), (2, 8, 16, 180, 540, 1),
    (1, 16, 16, 180, 180, 3), (3, 8, 8, 64, 256, 3), (2, 8, 8, 360, 180, 1), (1, 40, 24, 20, 70, 3),
    (2, 24, 40, 3, 64, 3), (2, 24, 40, 64, 3, 3), (1, 16, 16, 128, 3, 3), (2, 16, 24, 4, 48, 3)])
def test_conv_fprop_dgrad_wgrad(engine, B, H, W, cin, cout, k):
    from neosr_b200 import ops
    if engine == "tcgen05" and (min(cin, cout) < 16 or cin % 4 or cout % 4):
        pytest.skip("image-side / odd-channel convs stay on the exact-fp32 engine by design")
    x = rnd(B, cin, H, W, seed=1).requires_grad_(True)
    w = rnd(cout, cin, k, k, seed=2, scale=1 / math.sqrt(cin * k * k)).requires_grad_(True)
    b = rnd(cout, seed=3, scale=0.1).requires_grad_(True)
    y_ref = F.conv2d(x, w, b, 1, k // 2)
    dy = rnd(B, cout, H, W, seed=4)
    y_ref.backward(dy)
    pw = ops.PackedWeight(w.detach()).refresh()
    y = ops.conv_fprop(nhwc(x.detach()), pw, b.detach(), engine=engine)
    assert rel(nchw(y), y_ref.detach()) < 1e-4
    dx = ops.conv_fprop(nhwc(dy), pw, None, dgrad=True, engine=engine)
    assert rel(nchw(dx), x.grad) < 1e-4
    dw, db = torch.empty_like(w), torch.empty_like(b)
    ops.conv_wgrad(nhwc(x.detach()), nhwc(dy), dw, db, k, k, engine=engine)
    assert rel(dw, w.grad) < 1e-4
    assert rel(db, b.grad) < 1e-4


@pytest.mark.parametrize("engine", ENGINES)
def test_conv_epilogues(engine):
    from neosr_b200 import ops
    B, H, W, cin, cout = 2, 8, 16, 180, 360
    x, res = rnd(B, H, W, cin, seed=1), rnd(B, H, W, cout, seed=5)
    w, b = rnd(cout, cin, seed=2, scale=1 / math.sqrt(cin)), rnd(cout, seed=3, scale=0.1)
    pw = ops.PackedWeight(w).refresh()
    lin = F.linear(x, w, b)
    y, pre = ops.conv_fprop(x, pw, b, act="gelu", want_pre=True, engine=engine)
    assert rel(pre, lin) < 1e-4 and rel(y, F.gelu(lin)) < 1e-4
    y = ops.conv_fprop(x, pw, b, act="lrelu", act_slope=0.01, residual=res, engine=engine)
    assert rel(y, F.leaky_relu(lin, 0.01) + res) < 1e-4
    rs = torch.tensor([0.0, 1.25]).cuda()
    y = ops.conv_fprop(x, pw, b, row_scale=rs, residual=res, engine=engine)
    assert rel(y, lin * rs.view(B, 1, 1, 1) + res) < 1e-4
    # actgrad: v *= gelu'(aux)
    aux = rnd(B, H, W, cout, seed=7).requires_grad_(True)
    (F.gelu(aux)).sum().backward()
    y = ops.conv_fprop(x, pw, None, actgrad="gelu", aux=aux.detach(), engine=engine)
    assert rel(y, F.linear(x, w) * aux.grad) < 1e-4
    y = ops.conv_fprop(x, pw, None, actgrad="relu", aux=res, residual=res, engine=engine)
    assert rel(y, F.linear(x, w) * (res > 0).float() + res) < 1e-4
    # pre_mode = 1: y_pre holds gelu'(pre) so that the backward epilogue is a plain multiply (mulaux)
    h = lin.detach().clone().requires_grad_(True)
    F.gelu(h).sum().backward()
    y, dact = ops.conv_fprop(x, pw, b, act="gelu", want_pre=True, pre_is_actgrad=True, engine=engine)
    assert rel(y, F.gelu(lin)) < 1e-4 and rel(dact, h.grad) < 1e-4
    y = ops.conv_fprop(x, pw, None, actgrad="mulaux", aux=dact, engine=engine)
    assert rel(y, F.linear(x, w) * h.grad) < 1e-4


@pytest.mark.parametrize("M,K,N,k", [(131072, 180, 540, 1), (131072, 360, 180, 1), (32 * 64 * 64, 180, 180, 3),
                                     (8 * 128 * 128, 64, 256, 3), (4 * 32 * 32, 512, 512, 3), (1000, 64, 64, 1)])
def test_tcgen05_large_vs_torch(M, K, N, k):
    """Full-size contraction shapes of C3 on the tcgen05 engine vs cuBLAS/cuDNN fp32 (TF32 off)."""
    from neosr_b200 import ops
    if k == 1:
        B, H, W = 1, 1, M
    else:
        side = {32 * 64 * 64: (32, 64, 64), 8 * 128 * 128: (8, 128, 128), 4 * 32 * 32: (4, 32, 32)}[M]
        B, H, W = side
    x = rnd(B, H, W, K, seed=11)
    w = rnd(N, K, k, k, seed=12, scale=1 / math.sqrt(K * k * k))
    b = rnd(N, seed=13, scale=0.1)
    res = rnd(B, H, W, N, seed=14)
    pw = ops.PackedWeight(w).refresh()
    y = ops.conv_fprop(x, pw, b, residual=res, engine="tcgen05")
    ref = nhwc(F.conv2d(nchw(x), w, b, 1, k // 2)) + res
    assert rel(y, ref) < 2e-5
    y2 = ops.conv_fprop(x, pw, b, residual=res, engine="tcgen05")
    assert torch.equal(y, y2)  # run-to-run deterministic
    # wgrad: dw = dy^T x over all pixels (split-K, MN-major operands)
    dy = rnd(B, H, W, N, seed=15)
    wg = w.double().requires_grad_(True)  # fp64 reference: the reduction runs over up to 131072 pixels
    F.conv2d(nchw(x).double(), wg, None, 1, k // 2).backward(nchw(dy).double())
    dw, db = torch.empty_like(w), torch.empty_like(b)
    ops.conv_wgrad(x, dy, dw, db, k, k, engine="tcgen05")
    assert rel(dw.double(), wg.grad) < 3e-5
    assert rel(db.double(), dy.double().sum((0, 1, 2))) < 1e-5
    dw2 = torch.empty_like(w)
    ops.conv_wgrad(x, dy, dw2, None, k, k, engine="tcgen05")
    assert torch.equal(dw, dw2)


def test_layout_pixelshuffle_maxpool_bitexact():
    from neosr_b200 import ops
    x = rnd(3, 3, 20, 28, seed=1)
    sc, sh = torch.tensor([2.0, 0.5, 4.0]).cuda(), torch.tensor([-1.0, 0.25, 0.0]).cuda()
    y = ops.nchw_to_nhwc_affine(x, sc, sh)
    assert torch.equal(y, nhwc(x * sc.view(1, 3, 1, 1) + sh.view(1, 3, 1, 1)))
    x2 = rnd(2, 70, 12, 20, seed=2)
    assert torch.equal(ops.nchw_to_nhwc_affine(x2, None, None), nhwc(x2))
    assert torch.equal(ops.nhwc_to_nchw_affine(nhwc(x2), None, None), x2)
    assert torch.equal(ops.nhwc_to_nchw_affine(nhwc(x), sc, sh), x * sc.view(1, 3, 1, 1) + sh.view(1, 3, 1, 1))
    for r, c in ((2, 64), (3, 5), (4, 3)):
        z = rnd(2, c * r * r, 6, 10, seed=3)
        ref = F.pixel_shuffle(z, r)
        out = ops.pixel_shuffle(nhwc(z), r)
        assert torch.equal(nchw(out), ref)  # bit-exact index map
        assert torch.equal(ops.pixel_unshuffle(out, r), nhwc(z))
        assert torch.equal(nchw(ops.pixel_unshuffle(nhwc(ref), r)), F.pixel_unshuffle(ref, r))
    p = rnd(2, 16, 12, 20, seed=4).requires_grad_(True)
    act = F.relu(p)
    pooled = F.max_pool2d(act, 2, 2)
    assert torch.equal(nchw(ops.maxpool2(nhwc(act.detach()))), pooled.detach())
    dyp = rnd(*pooled.shape, seed=5)
    pooled.backward(dyp)
    extra = rnd(*p.shape, seed=6)
    dx = ops.maxpool2_relu_bwd(nhwc(act.detach()), nhwc(dyp), nhwc(extra))
    assert rel(nchw(dx), p.grad + extra) < 1e-6


@pytest.mark.parametrize("rows,c", [(64, 180), (1000, 36), (37, 360)])
def test_layernorm(rows, c):
    from neosr_b200 import ops
    x = rnd(rows, c, seed=1).requires_grad_(True)
    g, b = (1 + 0.1 * rnd(c, seed=2)).requires_grad_(True), rnd(c, seed=3, scale=0.1).requires_grad_(True)
    y_ref = F.layer_norm(x, (c,), g, b, 1e-5)
    dy, dres = rnd(rows, c, seed=4), rnd(rows, c, seed=5)
    y_ref.backward(dy)
    y, mu, rs = ops.layernorm_fwd(x.detach(), g.detach(), b.detach())
    assert rel(y, y_ref.detach()) < 1e-5
    dg, db = torch.empty_like(g), torch.empty_like(b)
    dx = ops.layernorm_bwd(dy, x.detach(), g.detach(), mu, rs, dg, db, dres=dres)
    assert rel(dx, x.grad + dres) < 1e-4
    assert rel(dg, g.grad) < 1e-4 and rel(db, b.grad) < 1e-4


@pytest.mark.parametrize("B,H,W,C,heads,ws,shift", [(2, 16, 16, 36, 3, 8, 0), (2, 16, 24, 36, 3, 8, 4),
                                                     (1, 64, 64, 180, 6, 8, 4), (1, 8, 16, 60, 6, 4, 2),
                                                     (1, 16, 16, 64, 2, 8, 4), (2, 8, 16, 96, 6, 8, 0)])  # head dim 32 (no padding), 16
def test_window_attention(B, H, W, C, heads, ws, shift):
    """vs the oracle's roll + window_partition + WindowAttention core + reverse (index-exact)."""
    from neosr_b200 import ops
    from oracle import swinir as O
    qkv = rnd(B, H, W, 3 * C, seed=1).requires_grad_(True)
    table = rnd((2 * ws - 1) ** 2, heads, seed=2, scale=0.5).requires_grad_(True)
    scale = (C // heads) ** -0.5
    N = ws * ws
    x = qkv
    if shift:
        x = torch.roll(x, (-shift, -shift), (1, 2))
    xw = O.window_partition(x, ws).view(-1, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = xw[0] * scale, xw[1], xw[2]
    attn = q @ k.transpose(-2, -1)
    idx = O.relative_position_index(ws).cuda()
    attn = attn + table[idx.view(-1)].view(N, N, -1).permute(2, 0, 1).unsqueeze(0)
    if shift:
        m = O.calculate_mask(H, W, ws, shift).cuda()
        attn = (attn.view(B, -1, heads, N, N) + m.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, ws, ws, C)
    o = O.window_reverse(o, ws, H, W)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    dout = rnd(B, H, W, C, seed=3)
    o.backward(dout)
    out = ops.window_attn_fwd(qkv.detach(), table.detach(), heads, ws, shift, scale)
    assert rel(out, o.detach()) < 5e-5
    dtable = torch.empty_like(table)
    dqkv = ops.window_attn_bwd(qkv.detach(), table.detach(), dout, dtable, heads, ws, shift, scale)
    assert rel(dqkv, qkv.grad) < 1e-4
    assert rel(dtable, table.grad) < 1e-4


def test_losses():
    from neosr_b200 import ops
    a = rnd(2, 3, 32, 48, seed=1).requires_grad_(True)
    b = rnd(2, 3, 32, 48, seed=2)
    acc = torch.zeros(1).cuda()
    ref = 0.7 * (a - b).abs().mean()
    ref.backward()
    v, g = ops.l1_loss(a.detach(), b, 0.7, acc)
    assert abs(float(v) - float(ref)) < 1e-6 and rel(g, a.grad) < 1e-6
    a.grad = None
    ref = 0.5 * torch.clamp(torch.sqrt((a / 10 - b / 10) ** 2 + 1e-12), 0, 0.2).mean()
    ref.backward()
    v, g = ops.charbonnier_loss(a.detach(), b, 0.5, acc, in_scale=0.1, clip_min=0.0, clip_max=0.2)
    assert abs(float(v) - float(ref)) < 1e-6 and rel(g, a.grad) < 1e-4
    a.grad = None
    ref = 0.3 * F.binary_cross_entropy_with_logits(a, torch.ones_like(a))
    ref.backward()
    v, g = ops.bce_logits_loss(a.detach(), 1.0, 0.3, acc)
    assert abs(float(v) - float(ref)) < 1e-6 and rel(g, a.grad) < 1e-5
    assert float(acc) == pytest.approx(float(0.7 * (a - b).abs().mean()) + float(
        0.5 * torch.clamp(torch.sqrt((a / 10 - b / 10) ** 2 + 1e-12), 0, 0.2).mean()) + float(ref), rel=1e-5)


def test_fused_adan_sf_clip_ema_vs_oracle():
    """3 steps of clip_grad_norm_(1.0) + adan_sf + EMA: fused kernel vs the oracle (CPU)."""
    from neosr_b200.optimizers import adan_sf
    from oracle.optim import AdanSFState, EMAState, adan_sf_step, clip_grad_norm
    shapes = [(180, 180, 3, 3), (540,), (225, 6), (5000,), (1,)]
    g0 = torch.Generator().manual_seed(0)
    cpu_p = [torch.randn(s, generator=g0) * 0.1 for s in shapes]
    kw = dict(lr=1e-3, betas=(0.98, 0.92, 0.987), weight_decay=0.02, schedule_free=True, warmup_steps=1600)
    st = AdanSFState([p.clone() for p in cpu_p], **kw)
    ema_o = EMAState(st.params, 0.999)
    params = [torch.nn.Parameter(p.clone().cuda()) for p in cpu_p]
    opt = adan_sf(params, **kw)
    ema = [p.detach().clone() for p in params]
    for it in range(3):
        grads = [torch.randn(s, generator=g0) * (3.0 if it == 0 else 0.01) for s in shapes]
        gc = [g.clone() for g in grads]
        clip_grad_norm(gc, 1.0)
        adan_sf_step(st, gc)
        ema_o.update(st.params)
        for p, g in zip(params, grads):
            p.grad = g.clone().cuda()
        opt.step(clip_max_norm=1.0, ema=(ema, 0.999, it == 0))
    for i in range(len(shapes)):
        assert rel(params[i].detach().cpu(), st.params[i]) < 1e-5
        assert rel(ema[i].cpu(), ema_o.avg[i]) < 1e-5
        s = opt.state[params[i]]
        assert rel(s["z"].cpu(), st.z[i]) < 1e-5
        assert rel(s["exp_avg_sq"].cpu(), st.exp_avg_sq[i]) < 1e-4
        assert rel(s["neg_pre_grad"].cpu(), st.neg_pre_grad[i]) < 1e-5
    g = opt.param_groups[0]
    assert g["step"] == 3 and abs(g["weight_sum"] - st.weight_sum) < 1e-15


# ----------------------------------------------------------------------------- split tile images
@pytest.mark.parametrize("B,H,W,C", [(1, 8, 24, 180), (2, 16, 16, 36), (1, 64, 64, 540)])
def test_sti_roundtrip(B, H, W, C):
    from neosr_b200 import ops
    x = rnd(B, H, W, C, seed=1)
    s = ops.STI.from_f32(x)
    assert rel(s.to_f32(), x) < 1e-5  # hi + lo carries ~17 mantissa bits


@pytest.mark.parametrize("B,H,W,cin,cout", [(1, 8, 24, 180, 540), (2, 32, 64, 180, 360), (2, 32, 64, 360, 180),
                                            (1, 16, 16, 36, 72), (32, 64, 64, 180, 180)])
def test_sti_linear_fprop_dgrad_wgrad(B, H, W, cin, cout):
    """1x1 contractions whose operands arrive / leave as split tile images (bulk-copy kernels)."""
    from neosr_b200 import ops
    x = rnd(B, H, W, cin, seed=1)
    w = rnd(cout, cin, seed=2, scale=1 / math.sqrt(cin))
    b = rnd(cout, seed=3, scale=0.1)
    res = rnd(B, H, W, cout, seed=4)
    dy = rnd(B, H, W, cout, seed=5)
    pw = ops.PackedWeight(w).refresh()
    xs, dys = ops.STI.from_f32(x), ops.STI.from_f32(dy)
    lin = F.linear(x.double(), w.double(), b.double())
    y = ops.conv_fprop(xs, pw, b, residual=res)
    assert rel(y.double(), lin + res.double()) < 2e-5
    (ys, pre) = ops.conv_fprop(xs, pw, b, act="gelu", want_pre=True, sti_out=True, f32_out=False)
    assert rel(pre.double(), lin) < 2e-5
    assert rel(ys.to_f32().double(), F.gelu(lin)) < 3e-5
    yf, ys2 = ops.conv_fprop(x, pw, b, sti_out=True)  # fp32-A kernel emitting both formats
    assert rel(ys2.to_f32(), yf) < 1e-5
    dx = ops.conv_fprop(dys, pw, None, dgrad=True)
    assert rel(dx.double(), dy.double() @ w.double()) < 2e-5
    aux = rnd(B, H, W, cin, seed=7)
    dxs = ops.conv_fprop(dys, pw, None, dgrad=True, actgrad="gelu", aux=aux, sti_out=True, f32_out=False)
    ag = aux.double().requires_grad_(True)
    F.gelu(ag).sum().backward()
    assert rel(dxs.to_f32().double(), (dy.double() @ w.double()) * ag.grad) < 3e-5
    dw, db = torch.empty_like(w), torch.empty_like(b)
    ops.conv_wgrad(None, None, dw, db, 1, 1, x_sti=xs, dy_sti=dys)
    ref_dw = dy.double().reshape(-1, cout).t() @ x.double().reshape(-1, cin)
    assert rel(dw.double(), ref_dw) < 3e-5
    assert rel(db.double(), dy.double().sum((0, 1, 2))) < 2e-5
    dw2, db2 = torch.empty_like(w), torch.empty_like(b)
    ops.conv_wgrad(None, dy, dw2, db2, 1, 1, x_sti=xs, dy_sti=dys)
    assert torch.equal(dw, dw2) and rel(db2.double(), dy.double().sum((0, 1, 2))) < 1e-5


def test_sti_layernorm_and_attention():
    from neosr_b200 import ops
    rows, c = 2 * 16 * 16, 180
    x = rnd(2, 16, 16, c, seed=1)
    g, b = 1 + 0.1 * rnd(c, seed=2), rnd(c, seed=3, scale=0.1)
    y_ref = F.layer_norm(x, (c,), g, b, 1e-5)
    (yf, ys), mu, rs = ops.layernorm_fwd(x, g, b, sti_out=True)
    assert rel(yf, y_ref) < 1e-5 and rel(ys.to_f32(), y_ref) < 2e-5
    dy, dres = rnd(2, 16, 16, c, seed=4), rnd(2, 16, 16, c, seed=5)
    dg, db = torch.empty_like(g), torch.empty_like(b)
    dx, dxs = ops.layernorm_bwd(dy, x, g, mu, rs, dg, db, dres=dres, sti_out=True)
    dx0 = ops.layernorm_bwd(dy, x, g, mu, rs, torch.empty_like(g), torch.empty_like(b), dres=dres)
    assert torch.equal(dx, dx0) and rel(dxs.to_f32(), dx) < 2e-5
    qkv = rnd(2, 16, 16, 3 * c, seed=6)
    table = rnd(225, 6, seed=7, scale=0.5)
    scale = 30 ** -0.5
    for shift in (0, 4):
        o = ops.window_attn_fwd(qkv, table, 6, 8, shift, scale)
        os_ = ops.window_attn_fwd(qkv, table, 6, 8, shift, scale, sti_out=True)
        assert rel(os_.to_f32(), o) < 2e-5
        dout = rnd(2, 16, 16, c, seed=8)
        dt1, dt2 = torch.empty_like(table), torch.empty_like(table)
        dq = ops.window_attn_bwd(qkv, table, dout, dt1, 6, 8, shift, scale)
        dqs = ops.window_attn_bwd(qkv, table, dout, dt2, 6, 8, shift, scale, sti_out=True)
        assert rel(dqs.to_f32(), dq) < 2e-5 and torch.equal(dt1, dt2)


@pytest.mark.parametrize("B,H,W,cin,cout,x_ld", [
    (2, 16, 16, 64, 32, 192), (2, 32, 32, 160, 32, 192), (1, 64, 64, 192, 64, 192), (2, 6, 48, 96, 32, 96),
    (1, 5, 128, 64, 64, 64), (3, 4, 16, 180, 180, 180), (2, 16, 32, 20, 72, 20), (1, 8, 64, 256, 128, 256)])
def test_wgrad_tma_3x3(B, H, W, cin, cout, x_ld):
    """TMA-staged 3x3 weight gradient (igemm_wgrad_tma.cu): halo tiles, OOB zero fill at the image border, slab
    (strided) inputs, every BN / orientation; vs fp64 autograd and vs the exact-fp32 engine."""
    from neosr_b200 import ops
    slab = rnd(B, H, W, x_ld, seed=21)
    x = ops.Slab(slab, 0, cin) if x_ld != cin else slab
    xd = slab[..., :cin].contiguous()
    dy = rnd(B, H, W, cout, seed=22)
    wg = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, device="cuda", requires_grad=True)
    F.conv2d(nchw(xd).double(), wg, None, 1, 1).backward(nchw(dy).double())
    dw, db = torch.empty(cout, cin, 3, 3, device="cuda"), torch.empty(cout, device="cuda")
    ops.conv_wgrad(x, dy, dw, db, 3, 3, engine="tcgen05")
    assert rel(dw.double(), wg.grad) < 3e-5
    assert rel(db.double(), dy.double().sum((0, 1, 2))) < 1e-5
    dw2 = torch.empty_like(dw)
    ops.conv_wgrad(x, dy, dw2, None, 3, 3, engine="tcgen05")
    assert torch.equal(dw, dw2)


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 96, 96, 3, 64), (2, 96, 96, 64, 3), (1, 128, 160, 1, 64), (1, 128, 160, 64, 1),
                                            (2, 96, 100, 4, 48), (3, 80, 72, 128, 3)])
def test_narrow_conv_as_gemm(B, H, W, cin, cout):
    """<= 4-channel convs at sizes where they run as im2col + tcgen05 contraction (conv_narrow_gemm.cu):
    fprop with epilogue, dgrad, wgrad + bias grad, vs torch fp32 autograd and vs the exact-fp32 engine."""
    from neosr_b200 import ops
    x = rnd(B, cin, H, W, seed=1).requires_grad_(True)
    w = rnd(cout, cin, 3, 3, seed=2, scale=1 / math.sqrt(cin * 9)).requires_grad_(True)
    b = rnd(cout, seed=3, scale=0.1).requires_grad_(True)
    y_ref = F.conv2d(x, w, b, 1, 1)
    dy = rnd(B, cout, H, W, seed=4)
    y_ref.backward(dy)
    pw = ops.PackedWeight(w.detach()).refresh()
    xh, dyh = nhwc(x.detach()), nhwc(dy)
    y = ops.conv_fprop(xh, pw, b.detach())
    assert rel(nchw(y), y_ref.detach()) < 2e-5
    assert rel(y, ops.conv_fprop(xh, pw, b.detach(), engine="simt")) < 2e-5
    if cin <= 4:  # fused epilogue of the wide side
        y2, pre = ops.conv_fprop(xh, pw, b.detach(), act="lrelu", act_slope=0.2, want_pre=True)
        assert rel(nchw(pre), y_ref.detach()) < 2e-5
        assert rel(nchw(y2), F.leaky_relu(y_ref.detach(), 0.2)) < 2e-5
    dx = ops.conv_fprop(dyh, pw, None, dgrad=True)
    assert rel(nchw(dx), x.grad) < 2e-5
    dw, db = torch.empty_like(w), torch.empty_like(b)
    ops.conv_wgrad(xh, dyh, dw, db, 3, 3)
    assert rel(dw, w.grad) < 5e-5
    assert rel(db, b.grad) < 1e-5


def test_bf16_single_pass_engine_contractions():
    """NSR_ENGINE_BF16 (`use_amp` + `bfloat16`): the same kernels with ONE bf16 pass - results within bf16 round-off of the
    fp64 product (a few 1e-3) and clearly NOT the three-pass result (which sits at 1e-5), on every tcgen05 contraction path:
    split-tile-image fprop / dgrad / wgrad, fp32-operand 3x3 fprop, TMA-staged 3x3 wgrad."""
    from neosr_b200 import ops
    B, H, W, cin, cout = 2, 32, 64, 180, 360
    x, dy = rnd(B, H, W, cin, seed=1), rnd(B, H, W, cout, seed=5)
    w, b = rnd(cout, cin, seed=2, scale=1 / math.sqrt(cin)), rnd(cout, seed=3, scale=0.1)
    pw = ops.PackedWeight(w).refresh()
    xs, dys = ops.STI.from_f32(x), ops.STI.from_f32(dy)
    lin = F.linear(x.double(), w.double(), b.double())
    ref_dw = dy.double().reshape(-1, cout).t() @ x.double().reshape(-1, cin)
    w3 = rnd(64, 64, 3, 3, seed=7, scale=1 / math.sqrt(576))
    pw3 = ops.PackedWeight(w3).refresh()
    x3, dy3 = rnd(2, 32, 32, 64, seed=8), rnd(2, 32, 32, 64, seed=9)
    conv3 = nhwc(F.conv2d(nchw(x3).double(), w3.double(), padding=1))
    xr = nchw(x3).double()
    wr = w3.double().requires_grad_(True)
    F.conv2d(xr, wr, padding=1).backward(nchw(dy3).double())
    errs = {}
    for engine in ("auto", "bf16"):
        ops.DEFAULT_ENGINE = engine
        try:
            e = {}
            e["sti fprop"] = rel(ops.conv_fprop(xs, pw, b).double(), lin)
            e["sti dgrad"] = rel(ops.conv_fprop(dys, pw, None, dgrad=True).double(), dy.double() @ w.double())
            dw = torch.empty_like(w)
            ops.conv_wgrad(None, None, dw, None, 1, 1, x_sti=xs, dy_sti=dys)
            e["sti wgrad"] = rel(dw.double(), ref_dw)
            e["3x3 fprop"] = rel(ops.conv_fprop(x3, pw3, None).double(), conv3)
            dw3 = torch.empty_like(w3)
            ops.conv_wgrad(x3, dy3, dw3, None, 3, 3)
            e["3x3 wgrad"] = rel(dw3.double(), wr.grad)
            errs[engine] = e
        finally:
            ops.DEFAULT_ENGINE = "auto"
    print(errs)
    for k in errs["auto"]:
        assert errs["auto"][k] < 5e-5, (k, errs["auto"][k])
        assert 2e-4 < errs["bf16"][k] < 2e-2, (k, errs["bf16"][k])
