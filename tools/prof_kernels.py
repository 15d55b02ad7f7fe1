"""Launch single hot kernels at their C3 shapes (for ncu captures and micro-timing).

    python tools/prof_kernels.py linear_qkv [reps]
"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from neosr_b200 import ops  # noqa: E402


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "linear_qkv"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    B, H, W, C = 32, 64, 64, 180
    if what.startswith("linear_") or what.startswith("wgrad_") or what.startswith("dgrad_"):
        kind, name = what.split("_", 1)
        cin, cout, k, extra = {"qkv": (180, 540, 1, {}), "proj": (180, 180, 1, {"res": 1}),
                               "fc1": (180, 360, 1, {"gelu": 1}), "fc2": (360, 180, 1, {"res": 1}),
                               "conv": (180, 180, 3, {"res": 1})}[name]
        x = rnd(B, H, W, cin, seed=1)
        w = rnd(cout, cin, k, k, seed=2, scale=1 / math.sqrt(cin * k * k))
        b = rnd(cout, seed=3, scale=0.1)
        res = rnd(B, H, W, cout, seed=4)
        pw = ops.PackedWeight(w).refresh()
        dy = rnd(B, H, W, cout, seed=5)
        dw, db = torch.empty_like(w), torch.empty_like(b)

        def run():
            if kind == "linear":
                if extra.get("gelu"):
                    ops.conv_fprop(x, pw, b, act="gelu", want_pre=True)
                else:
                    ops.conv_fprop(x, pw, b, residual=res if extra.get("res") else None)
            elif kind == "dgrad":
                ops.conv_fprop(dy, pw, None, dgrad=True, actgrad="gelu" if name == "fc2" else "none",
                               aux=x if name == "fc2" else None)
            else:
                ops.conv_wgrad(x, dy, dw, db, k, k)
    elif what == "vgg_conv1_2":
        # the largest 3x3 producer-warp contraction of C3: 64 -> 64 channels at 32 x 256 x 256 pixels, ReLU epilogue
        x = rnd(32, 256, 256, 64, seed=1)
        w = rnd(64, 64, 3, 3, seed=2, scale=1 / math.sqrt(576))
        b = rnd(64, seed=3, scale=0.1)
        pw = ops.PackedWeight(w).refresh()

        def run():
            ops.conv_fprop(x, pw, b, act="relu")
    elif what.startswith("sti"):
        # STI-operand 1x1 contractions: stifprop_qkv, stidgrad_fc2, stiwgrad_qkv ...
        kind, name = what[3:].split("_", 1)
        cin, cout = {"qkv": (180, 540), "proj": (180, 180), "fc1": (180, 360), "fc2": (360, 180)}[name]
        x = rnd(B, H, W, cin, seed=1)
        w = rnd(cout, cin, seed=2, scale=1 / math.sqrt(cin))
        b = rnd(cout, seed=3, scale=0.1)
        res = rnd(B, H, W, cout, seed=4)
        dy = rnd(B, H, W, cout, seed=5)
        pw = ops.PackedWeight(w).refresh()
        xs, dys = ops.STI.from_f32(x), ops.STI.from_f32(dy)
        dw, db = torch.empty_like(w), torch.empty_like(b)

        def run():
            if kind == "fprop":
                if name == "fc1":
                    ops.conv_fprop(xs, pw, b, act="gelu", want_pre=True, pre_is_actgrad=True, sti_out=True, f32_out=False)
                else:
                    ops.conv_fprop(xs, pw, b, residual=res if name in ("proj", "fc2") else None)
            elif kind == "dgrad":
                if name == "fc2":
                    ops.conv_fprop(dys, pw, None, dgrad=True, actgrad="gelu", aux=x, sti_out=True, f32_out=False)
                else:
                    ops.conv_fprop(dys, pw, None, dgrad=True)
            else:
                ops.conv_wgrad(None, dy, dw, db, 1, 1, x_sti=xs, dy_sti=dys)
    elif what.startswith("ln"):
        x = rnd(B, H, W, C, seed=1)
        g, bb = 1 + 0.1 * rnd(C, seed=2), rnd(C, seed=3, scale=0.1)
        dy, dres = rnd(B, H, W, C, seed=4), rnd(B, H, W, C, seed=5)
        _, mu, rs = ops.layernorm_fwd(x, g, bb)
        dg, db = torch.empty_like(g), torch.empty_like(bb)

        def run():
            if what == "ln_fwd":
                ops.layernorm_fwd(x, g, bb, sti_out=True, f32_out=False)
            else:
                ops.layernorm_bwd(dy, x, g, mu, rs, dg, db, dres=dres, sti_out=True)
    elif what.startswith("attn"):
        qkv = rnd(B, H, W, 3 * C, seed=1)
        table = rnd(225, 6, seed=2, scale=0.5)
        dout = rnd(B, H, W, C, seed=3)
        dtable = torch.empty_like(table)

        def run():
            if what == "attn_fwd":
                ops.window_attn_fwd(qkv, table, 6, 8, 4, 30 ** -0.5, sti_out=True)
            else:
                ops.window_attn_bwd(qkv, table, dout, dtable, 6, 8, 4, 30 ** -0.5, sti_out=True)
    else:
        raise SystemExit(f"unknown kernel {what}")
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"{what}: {e0.elapsed_time(e1) / reps:.4f} ms per call over {reps} reps")


if __name__ == "__main__":
    main()
