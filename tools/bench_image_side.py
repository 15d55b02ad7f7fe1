"""Image-side (<= 4 channels on one side) 3x3 convolutions at the C3 output resolution (32 x 256 x 256), L2 flushed between
launches:  python tools/bench_image_side.py
(Round 2 tried a row-sliding direct kernel for the 64 -> 3 weight gradient - 3 x 3 window of dy in registers, x rows
prefetched four deep: 1185 us against 1065 us for the im2col + tcgen05 route at 160 registers / 12 warps per SM; not kept.)"""
import math
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from neosr_b200 import ops  # noqa: E402


def timed(fn, flush, iters=10):
    fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * iters)]
    for i in range(iters):
        flush.zero_()
        ev[2 * i].record()
        fn()
        ev[2 * i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) * 1e3 for i in range(iters))
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(1)
    B, H, W = 32, 256, 256
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    x64 = torch.randn(B, H, W, 64, generator=g, device=dev)
    x3 = torch.randn(B, H, W, 3, generator=g, device=dev)
    w_last = torch.randn(3, 64, 3, 3, generator=g, device=dev) / math.sqrt(576)   # conv_last: 64 -> 3
    w_first = torch.randn(64, 3, 3, 3, generator=g, device=dev) / math.sqrt(27)   # VGG conv1_1: 3 -> 64
    pw_last, pw_first = ops.PackedWeight(w_last).refresh(), ops.PackedWeight(w_first).refresh()
    dw_last, db_last = torch.empty_like(w_last), torch.empty(3, device=dev)
    dw_first, db_first = torch.empty_like(w_first), torch.empty(64, device=dev)
    # correctness of the wide -> narrow weight gradient on a slice (fp64 reference)
    xs, ds = x64[:2, :48, :40].contiguous(), x3[:2, :48, :40].contiguous()
    dws = torch.empty_like(w_last)
    ops.conv_wgrad(xs, ds, dws, None, 3, 3)
    wr = w_last.double().requires_grad_(True)
    F.conv2d(xs.permute(0, 3, 1, 2).double(), wr, padding=1).backward(ds.permute(0, 3, 1, 2).double())
    err = ((dws.double() - wr.grad).abs().max() / wr.grad.abs().max()).item()
    print(f"wgrad 64->3 rel err {err:.2e}")
    assert err < 1e-5
    for name, fn in (
        ("conv_last  fprop 64->3", lambda: ops.conv_fprop(x64, pw_last, None)),
        ("conv_last  dgrad 3->64", lambda: ops.conv_fprop(x3, pw_last, None, dgrad=True)),
        ("conv_last  wgrad", lambda: ops.conv_wgrad(x64, x3, dw_last, db_last, 3, 3)),
        ("vgg conv1_1 fprop 3->64", lambda: ops.conv_fprop(x3, pw_first, None, act="relu")),
        ("vgg conv1_1 dgrad 64->3", lambda: ops.conv_fprop(x64, pw_first, None, dgrad=True)),
        ("vgg conv1_1 wgrad", lambda: ops.conv_wgrad(x3, x64, dw_first, db_first, 3, 3)),
    ):
        print(f"{name:26s} {timed(fn, flush):8.1f} us")


if __name__ == "__main__":
    main()
