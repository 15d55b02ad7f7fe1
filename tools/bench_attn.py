"""Micro-benchmark of the window-attention kernels at the C3 shape (B=32, 64x64 tokens, C=180, 6 heads):
    python tools/bench_attn.py [--iters 20] [--only fwd_tc|fwd_mma|bwd_tc|bwd_mma] [--shift 4]
CUDA-event timing per engine; operands (302 MB qkv image) exceed L2, so every iteration reads HBM."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from neosr_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--shift", type=int, default=4)
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    B, H, W, C, heads, ws = a.batch, 64, 64, 180, 6, 8
    scale = (C // heads) ** -0.5
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, H, W, C, generator=g).cuda()
    wq = (torch.randn(3 * C, C, generator=g) * C ** -0.5).cuda()
    bq = torch.zeros(3 * C).cuda()
    table = (torch.randn(225, heads, generator=g) * 0.5).cuda()
    dout = torch.randn(B, H, W, C, generator=g).cuda()
    xs = ops.STI.from_f32(x)
    qw = ops.MappedPackedWeight(wq, bq, row_map=ops.head_pad_map(C, heads, 3), need_dgrad=False).refresh()
    qkv = ops.conv_fprop(xs, qw, qw.bias_padded, sti_out=True, f32_out=False, sti_win=(ws, a.shift))
    pw = ops.MappedPackedWeight(torch.eye(C, device="cuda"), None, col_map=ops.head_pad_map(C, heads, 1)).refresh()
    dsti = ops.conv_fprop(ops.STI.from_f32(dout), pw, None, dgrad=True, sti_out=True, f32_out=False, sti_win=(ws, a.shift))
    dtab = torch.zeros_like(table)
    cases = {
        "fwd_tc": lambda: ops.window_attn_fwd_wsti(qkv, table, C, heads, ws, a.shift, scale, engine="tcgen05", padded_out=True),
        "fwd_mma": lambda: ops.window_attn_fwd_wsti(qkv, table, C, heads, ws, a.shift, scale, engine="mma_sync"),
        "bwd_tc": lambda: ops.window_attn_bwd_wsti(qkv, table, dsti, dtab, C, heads, ws, a.shift, scale, engine="tcgen05", padded_out=True),
        "bwd_mma": lambda: ops.window_attn_bwd_wsti(qkv, table, dsti, dtab, C, heads, ws, a.shift, scale, engine="mma_sync"),
    }
    tokens = B * H * W
    for name, fn in cases.items():
        if a.only and name != a.only:
            continue
        try:
            for _ in range(3):
                fn()
        except Exception as e:  # engine not built
            print(f"{name}: skipped ({e})")
            continue
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / a.iters * 1e3
        nbytes = tokens * 4 * ((3 * 192 + 192) if name.startswith("fwd") else (3 * 192 + 192 + 576))
        print(f"{name}: {us:8.1f} us  {nbytes / us / 1e3:7.0f} GB/s (image bytes)")


if __name__ == "__main__":
    main()
