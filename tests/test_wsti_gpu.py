"""Window-ordered attention operands (WSTI): the qkv contraction writes a split tile image whose rows are in window order
(roll + window_partition folded into the epilogue store) and whose heads are padded to 32 channels; the attention
kernels bulk-copy it.  Checked against the token-order fp32 path of the same library (itself checked against the
PyTorch restatement of swinir_arch.py:150-212 in test_kernels_gpu.py) and directly against that restatement."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [dict(B=2, H=16, W=24, C=36, heads=3, shift=0), dict(B=2, H=16, W=24, C=36, heads=3, shift=4),
         dict(B=1, H=8, W=24, C=24, heads=2, shift=4),          # 3 windows: the last 128-row block is half empty
         dict(B=1, H=64, W=64, C=180, heads=6, shift=4), dict(B=2, H=32, W=32, C=180, heads=6, shift=0)]


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _torch_attention(qkv, table, heads, ws, shift):
    """swinir_arch.py:150-212 + 353-386 on a token-order qkv tensor [B,H,W,3C] (fp64)."""
    from oracle.swinir import calculate_mask, relative_position_index, window_partition, window_reverse
    B, H, W, c3 = qkv.shape
    c = c3 // 3
    x = torch.roll(qkv, shifts=(-shift, -shift), dims=(1, 2)) if shift else qkv
    xw = window_partition(x, ws).view(-1, ws * ws, 3, heads, c // heads).permute(2, 0, 3, 1, 4)
    q, k, v = xw[0] * (c // heads) ** -0.5, xw[1], xw[2]
    attn = q @ k.transpose(-2, -1)
    bias = table[relative_position_index(ws).view(-1)].view(ws * ws, ws * ws, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if shift:
        mask = calculate_mask(H, W, ws, shift).to(qkv.dtype)
        nw = mask.shape[0]
        attn = (attn.view(-1, nw, heads, ws * ws, ws * ws) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, ws * ws, ws * ws)
    o = (torch.softmax(attn, -1) @ v).transpose(1, 2).reshape(-1, ws, ws, c)
    o = window_reverse(o, ws, H, W)
    return torch.roll(o, shifts=(shift, shift), dims=(1, 2)) if shift else o


@pytest.mark.parametrize("engine", ["tcgen05", "mma_sync"])
@pytest.mark.parametrize("case", CASES)
def test_wsti_attention_forward_backward(case, engine):
    from neosr_b200 import ops
    B, H, W, C, heads, shift = (case[k] for k in ("B", "H", "W", "C", "heads", "shift"))
    ws, scale = 8, (C // heads) ** -0.5
    if not ops.wsti_supported(C, heads, ws):
        pytest.skip("needs the tcgen05 engine")
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, H, W, C, generator=g).cuda()
    wq = (torch.randn(3 * C, C, generator=g) * C ** -0.5).cuda()
    bq = (torch.randn(3 * C, generator=g) * 0.1).cuda()
    table = (torch.randn((2 * ws - 1) ** 2, heads, generator=g) * 0.5).cuda()
    dout = torch.randn(B, H, W, C, generator=g).cuda()
    xs = ops.STI.from_f32(x)

    # token-order fp32 path of the library
    qkv = ops.conv_fprop(xs, ops.PackedWeight(wq).refresh(), bq)
    att = ops.window_attn_fwd(qkv, table, heads, ws, shift, scale)
    dtab = torch.zeros_like(table)
    dqkv = ops.window_attn_bwd(qkv, table, dout, dtab, heads, ws, shift, scale)

    # window-ordered path: head-padded weights, sti_win epilogue, bulk-copied operands
    qw = ops.MappedPackedWeight(wq, bq, row_map=ops.head_pad_map(C, heads, 3), need_dgrad=False).refresh()
    G = len(ops.head_pad_map(C, heads, 1))
    qkv_w = ops.conv_fprop(xs, qw, qw.bias_padded, sti_out=True, f32_out=False, sti_win=(ws, shift))
    assert qkv_w.shape == (B, H, W, 3 * G)
    att_w = ops.window_attn_fwd_wsti(qkv_w, table, C, heads, ws, shift, scale, sti_out=False, engine=engine)
    att_s = ops.window_attn_fwd_wsti(qkv_w, table, C, heads, ws, shift, scale, sti_out=True, engine=engine)
    eye = torch.eye(C, device="cuda")  # proj = identity: its dgrad hands dout through, re-ordered and head-padded
    pw = ops.MappedPackedWeight(eye, None, col_map=ops.head_pad_map(C, heads, 1)).refresh()
    dout_w = ops.conv_fprop(ops.STI.from_f32(dout), pw, None, dgrad=True, sti_out=True, f32_out=False, sti_win=(ws, shift))
    assert dout_w.shape == (B, H, W, G)
    dtab_w = torch.zeros_like(table)
    dqkv_w = ops.window_attn_bwd_wsti(qkv_w, table, dout_w, dtab_w, C, heads, ws, shift, scale, sti_out=False, engine=engine)
    if engine == "tcgen05":  # head-padded dqkv image [.., 3G]: real channels at the padded slots, zeros elsewhere
        dq_p = ops.window_attn_bwd_wsti(qkv_w, table, dout_w, torch.zeros_like(table), C, heads, ws, shift, scale, sti_out=True,
                                        engine=engine, padded_out=True)
        assert dq_p.shape == (B, H, W, 3 * G)
        full = dq_p.to_f32()
        m3 = torch.tensor(ops.head_pad_map(C, heads, 3), device="cuda")
        assert float(full[..., m3 < 0].abs().max()) == 0.0
        dqkv_s_f32 = full[..., m3 >= 0]
        # the qkv contraction's gradients from the padded image: wgrad (+ bias from ln1's ones column) and dgrad
        dwq = torch.zeros_like(wq)  # (bias gradient: needs LayerNorm's ones column in x; covered by the whole-network tests)
        ops.conv_wgrad_mapped_rows(xs, dq_p, dwq, None, qw.row_map)
        qw2 = ops.MappedPackedWeight(wq, bq, row_map=ops.head_pad_map(C, heads, 3)).refresh()
        dx_p = ops.conv_fprop(dq_p, qw2, None, dgrad=True)
        dq_c = ops.STI.from_f32(dqkv_s_f32.contiguous())
        dwr = torch.zeros_like(wq)
        ops.conv_wgrad(None, None, dwr, None, 1, 1, x_sti=xs, dy_sti=dq_c)
        dx_r = ops.conv_fprop(dq_c, ops.PackedWeight(wq).refresh(), None, dgrad=True)
        assert rel(dwq, dwr) < 2e-5 and rel(dx_p, dx_r) < 2e-5
    else:
        dqkv_s_f32 = ops.window_attn_bwd_wsti(qkv_w, table, dout_w, torch.zeros_like(table), C, heads, ws, shift, scale,
                                              sti_out=True, engine=engine).to_f32()

    assert rel(att_w, att) < 1e-4
    assert rel(att_s.to_f32(), att) < 1e-4
    if engine == "tcgen05":  # head-padded output image: heads at 32-channel slots, 1.0 in head 0's first padding slot
        att_p = ops.window_attn_fwd_wsti(qkv_w, table, C, heads, ws, shift, scale, sti_out=True, engine=engine, padded_out=True)
        assert att_p.shape == (B, H, W, G) and att_p.ones_col == C // heads
        full = att_p.to_f32()
        m = torch.tensor(ops.head_pad_map(C, heads, 1), device="cuda")
        assert rel(full[..., m >= 0], att) < 1e-4
        pad = full[..., m < 0]
        assert torch.equal(pad[..., 0], torch.ones_like(pad[..., 0])) and float(pad[..., 1:].abs().max()) == 0.0
        # proj on the padded image: fprop over G channels with zero weight columns, wgrad un-padded, dbias from the ones column
        wp = (torch.randn(C, C, generator=torch.Generator().manual_seed(13)) * C ** -0.5).cuda()
        bp = torch.randn(C, generator=torch.Generator().manual_seed(14)).cuda()
        pwm = ops.MappedPackedWeight(wp, None, col_map=ops.head_pad_map(C, heads, 1)).refresh()
        y_pad = ops.conv_fprop(att_p, pwm, bp)
        y_ref = ops.conv_fprop(att_s, ops.PackedWeight(wp).refresh(), bp)
        assert rel(y_pad, y_ref) < 2e-5
        gsti = ops.STI.from_f32(dout)
        dw, db = torch.zeros_like(wp), torch.zeros_like(bp)
        ops.conv_wgrad_mapped(att_p, gsti, dw, db, pwm.col_map, att_p.ones_col)
        dw_ref, db_ref = torch.zeros_like(wp), torch.zeros_like(bp)
        ops.conv_wgrad(None, None, dw_ref, db_ref, 1, 1, x_sti=att_s, dy_sti=gsti)
        assert rel(dw, dw_ref) < 2e-5 and rel(db, db_ref) < 2e-5
    if C % 64:  # first padding channel of the image = 1.0 (bias-gradient column of proj's wgrad), the rest 0
        kb = (C + 63) // 64 * 64
        raw = torch.empty(B, H, W, kb, device="cuda")
        from neosr_b200 import _lib
        _lib.check(_lib.lib().nsr_sti_to_f32(att_s.data_ptr(), B * H * W, kb, raw.data_ptr(), kb, ops._stream()), "sti_to_f32")
        assert torch.equal(raw[..., C], torch.ones_like(raw[..., C])) and float(raw[..., C + 1:].abs().max()) == 0.0
    assert rel(dqkv_w, dqkv) < 2e-4
    assert rel(dqkv_s_f32, dqkv) < 2e-4
    assert rel(dtab_w, dtab) < 2e-4
    # and against the PyTorch restatement in fp64
    q64 = (x.double().cpu().view(-1, C) @ wq.double().cpu().t() + bq.double().cpu()).view(B, H, W, 3 * C).requires_grad_(True)
    t64 = table.double().cpu().requires_grad_(True)
    ref = _torch_attention(q64, t64, heads, ws, shift)
    gq, gt = torch.autograd.grad(ref, [q64, t64], grad_outputs=dout.double().cpu())
    assert rel(att_w, ref) < 1e-4
    assert rel(dqkv_w, gq) < 2e-4
    assert rel(dtab_w, gt) < 2e-4
    # run-to-run determinism (fixed-order reductions)
    dtab_2 = torch.zeros_like(table)
    dqkv_2 = ops.window_attn_bwd_wsti(qkv_w, table, dout_w, dtab_2, C, heads, ws, shift, scale, sti_out=False, engine=engine)
    assert torch.equal(dqkv_2, dqkv_w) and torch.equal(dtab_2, dtab_w)


def test_window_ordered_sti_row_and_channel_layout():
    """The sti_win store itself: every (token, channel) of the padded qkv lands where roll + window_partition + the head
    padding say (bit-exact against the token-order image of the same contraction, re-indexed on the host)."""
    from neosr_b200 import ops
    B, H, W, C, heads, ws, shift = 2, 16, 24, 36, 3, 8, 4
    if not ops.wsti_supported(C, heads, ws):
        pytest.skip("needs the tcgen05 engine")
    g = torch.Generator().manual_seed(12)
    x = torch.randn(B, H, W, C, generator=g).cuda()
    wq = (torch.randn(3 * C, C, generator=g) * C ** -0.5).cuda()
    bq = torch.randn(3 * C, generator=g).cuda()
    xs = ops.STI.from_f32(x)
    qw = ops.MappedPackedWeight(wq, bq, row_map=ops.head_pad_map(C, heads, 3), need_dgrad=False).refresh()
    tok = ops.conv_fprop(xs, qw, qw.bias_padded, sti_out=True, f32_out=False).to_f32()            # token order
    win = ops.conv_fprop(xs, qw, qw.bias_padded, sti_out=True, f32_out=False, sti_win=(ws, shift)).to_f32()
    from oracle.swinir import window_partition
    want = window_partition(torch.roll(tok, shifts=(-shift, -shift), dims=(1, 2)), ws).reshape(B, H, W, -1)
    assert torch.equal(win.reshape(-1, win.shape[-1]), want.reshape(-1, want.shape[-1]))
    m = torch.tensor(ops.head_pad_map(C, heads, 3))
    assert torch.equal(tok[..., m < 0], torch.zeros_like(tok[..., m < 0]))  # head padding is exactly zero


def test_deferred_reductions_match_the_immediate_ones():
    """LayerNorm dgamma / dbeta and the attention bias-table gradient reduced LATER, batched over layers
    (`DeferredWgrads.finalize`: nsr_wgrad_finalize_multi + nsr_window_attn_dbias_multi), against the per-layer reductions of
    the same partial sums; two layers of different size in one batch."""
    from neosr_b200 import ops
    ws = 8
    g = torch.Generator().manual_seed(5)
    deferred = ops.DeferredWgrads()
    want = []
    for li, (B, H, W, C, heads, shift) in enumerate([(2, 16, 24, 36, 3, 4), (1, 32, 32, 180, 6, 0)]):
        if not ops.wsti_supported(C, heads, ws):
            pytest.skip("needs the tcgen05 engine")
        scale = (C // heads) ** -0.5
        x = (torch.randn(B, H, W, C, generator=g) * 1.5 + 0.2).cuda()
        gm, bt = (torch.randn(C, generator=g) * 0.2 + 1).cuda(), (torch.randn(C, generator=g) * 0.1).cuda()
        dy, dres = torch.randn(B, H, W, C, generator=g).cuda(), torch.randn(B, H, W, C, generator=g).cuda()
        _, mu, rs = ops.layernorm_fwd(x, gm, bt)
        dg0, db0 = torch.empty_like(gm), torch.empty_like(bt)
        dx0 = ops.layernorm_bwd(dy, x, gm, mu, rs, dg0, db0, dres=dres)
        dg1, db1 = torch.full_like(gm, float("nan")), torch.full_like(bt, float("nan"))
        dx1 = ops.layernorm_bwd(dy, x, gm, mu, rs, dg1, db1, dres=dres, deferred=deferred, key=f"ln{li}")
        assert torch.equal(dx0, dx1)
        # attention: window-ordered operands as in test_wsti_attention_forward_backward
        wq = (torch.randn(3 * C, C, generator=g) * C ** -0.5).cuda()
        bq = (torch.randn(3 * C, generator=g) * 0.1).cuda()
        table = (torch.randn((2 * ws - 1) ** 2, heads, generator=g) * 0.5).cuda()
        dout = torch.randn(B, H, W, C, generator=g).cuda()
        qw = ops.MappedPackedWeight(wq, bq, row_map=ops.head_pad_map(C, heads, 3), need_dgrad=False).refresh()
        qkv_w = ops.conv_fprop(ops.STI.from_f32(x), qw, qw.bias_padded, sti_out=True, f32_out=False, sti_win=(ws, shift))
        pw = ops.MappedPackedWeight(torch.eye(C, device="cuda"), None, col_map=ops.head_pad_map(C, heads, 1)).refresh()
        dout_w = ops.conv_fprop(ops.STI.from_f32(dout), pw, None, dgrad=True, sti_out=True, f32_out=False, sti_win=(ws, shift))
        dt0 = torch.zeros_like(table)
        dq0 = ops.window_attn_bwd_wsti(qkv_w, table, dout_w, dt0, C, heads, ws, shift, scale, sti_out=True, padded_out=True)
        dt1 = torch.full_like(table, float("nan"))
        dq1 = ops.window_attn_bwd_wsti(qkv_w, table, dout_w, dt1, C, heads, ws, shift, scale, sti_out=True, padded_out=True,
                                       deferred=deferred, key=f"attn{li}")
        assert torch.equal(dq0.to_f32(), dq1.to_f32())
        want.append((dg0, db0, dt0, dg1, db1, dt1))
    assert all(bool(torch.isnan(t).all()) for w in want for t in w[3:])  # nothing written before the batched reduction
    deferred.finalize()
    for dg0, db0, dt0, dg1, db1, dt1 in want:
        assert rel(dg1, dg0) < 1e-5 and rel(db1, db0) < 1e-5  # same partial sums, different (fixed) summation order
        assert torch.equal(dt1, dt0)  # same two-step reduction, batched
