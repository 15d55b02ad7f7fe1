"""Times the eight 1x1 contractions (forward + dgrad) of a SwinIR-medium block at the C3 token count on split-tile-image
operands, with the epilogues the engine uses, L2 flushed between launches:
    python tools/bench_sti_fprop.py [--iters 20]"""
import argparse
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from neosr_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(1)
    B, H, W = 32, 64, 64

    def rnd(*s, scale=1.0):
        return torch.randn(*s, generator=g, device=dev) * scale

    def pw(cout, cin):
        return ops.PackedWeight(rnd(cout, cin, 1, 1, scale=1 / math.sqrt(cin))).refresh()

    x180, x192, x360, x576 = (ops.STI.from_f32(rnd(B, H, W, c)) for c in (180, 192, 360, 576))
    res = rnd(B, H, W, 180)
    w_qkv, w_proj, w_fc1, w_fc2 = pw(576, 180), pw(180, 192), pw(360, 180), pw(180, 360)
    b576, b180, b360 = rnd(576), rnd(180), rnd(360)
    _, hpre = ops.conv_fprop(x180, w_fc1, b360, act="gelu", want_pre=True, pre_is_actgrad=True, sti_out=True, f32_out=False,
                             pre_u16=ops.AGC_U16)
    cases = [
        ("qkv    180->576 win-STI", lambda: ops.conv_fprop(x180, w_qkv, b576, sti_out=True, f32_out=False, sti_win=(8, 4))),
        ("proj   192->180 +res", lambda: ops.conv_fprop(x192, w_proj, b180, residual=res)),
        ("fc1    180->360 gelu", lambda: ops.conv_fprop(x180, w_fc1, b360, act="gelu", want_pre=True, pre_is_actgrad=True,
                                                        sti_out=True, f32_out=False, pre_u16=ops.AGC_U16)),
        ("fc2    360->180 +res", lambda: ops.conv_fprop(x360, w_fc2, b180, residual=res)),
        ("fc2^T  180->360 *aux", lambda: ops.conv_fprop(x180, w_fc2, None, dgrad=True, actgrad="mulaux", aux=hpre, sti_out=True,
                                                        f32_out=False)),
        ("fc1^T  360->180", lambda: ops.conv_fprop(x360, w_fc1, None, dgrad=True)),
        ("proj^T 180->192 win-STI", lambda: ops.conv_fprop(x180, w_proj, None, dgrad=True, sti_out=True, f32_out=False,
                                                           sti_win=(8, 4))),
        ("qkv^T  576->180", lambda: ops.conv_fprop(x576, w_qkv, None, dgrad=True)),
    ]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    total = 0.0
    for name, fn in cases:
        fn()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * a.iters)]
        for i in range(a.iters):
            flush.zero_()
            ev[2 * i].record()
            fn()
            ev[2 * i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) * 1e3 for i in range(a.iters))
        total += ts[len(ts) // 2]
        print(f"{name:26s} median {ts[len(ts) // 2]:7.1f} us   min {ts[0]:7.1f} us")
    print(f"sum of medians {total:7.1f} us per block")


if __name__ == "__main__":
    main()
