"""Times the four 1x1 weight-gradient shapes of a SwinIR-medium block at the C3 token count on split-tile-image operands
(`nsr_conv_wgrad` with x_sti / dy_sti) and checks each against a float64 reference:
    python tools/bench_wgrad.py [--rows 131072] [--iters 20]
Process-wide knobs of the kernel (read once): NSR_WG_STI2=0 (previous 2-stage kernel), NSR_WG_KPIX=32|64, NSR_WG_PAIR=0|1."""
import argparse
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from neosr_b200 import ops  # noqa: E402

SHAPES = [("qkv", 184, 576), ("proj", 192, 180), ("fc1", 184, 360), ("fc2", 364, 180)]  # (name, cin incl. ones col, cout)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=131072)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(1)
    B, H, W = a.rows // 4096, 64, 64
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    knobs = {k: os.environ.get(k, "-") for k in ("NSR_WG_STI2", "NSR_WG_KPIX", "NSR_WG_PAIR")}
    print("knobs", knobs)
    for name, cin, cout in SHAPES:
        x = torch.randn(B, H, W, cin, generator=g, device=dev)
        dy = torch.randn(B, H, W, cout, generator=g, device=dev) * 0.1
        xs, ds = ops.STI.from_f32(x), ops.STI.from_f32(dy)
        dw = torch.empty(cout, cin, 1, 1, device=dev)
        ops.conv_wgrad(None, None, dw, None, 1, 1, x_sti=xs, dy_sti=ds)
        ref = (dy.reshape(-1, cout).double().t() @ x.reshape(-1, cin).double())
        err = ((dw.reshape(cout, cin).double() - ref).abs().max() / ref.abs().max()).item()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * a.iters)]
        for i in range(a.iters):
            flush.zero_()
            ev[2 * i].record()
            ops.conv_wgrad(None, None, dw, None, 1, 1, x_sti=xs, dy_sti=ds)
            ev[2 * i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) * 1e3 for i in range(a.iters))
        mb = a.rows * 4 * (((cin + 63) // 64) * 64 + ((cout + 63) // 64) * 64) / 1e6
        med = ts[len(ts) // 2]
        print(f"{name:5s} cin {cin:3d} cout {cout:3d}: median {med:7.1f} us  min {ts[0]:7.1f} us  (incl. reduce)  "
              f"{mb / med * 1e-3:5.2f} TB/s of operand bytes  max rel err {err:.2e}")
        assert err < 1e-4, err


if __name__ == "__main__":
    main()
