"""LayerNorm forward / backward at the C3 token count (131072 x 180), split-tile-image output as in the Swin blocks, L2
flushed between launches; checks against torch:  python tools/bench_ln.py   (knob: NSR_LN_V3=0|1|2 = rows per half-warp of the forward kernel, 0 = previous kernel)
Measured: the tile-image-residual backward (101 us) is occupancy-bound - three blocks per SM without its 36 bytes of spills
ran 134 us."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from neosr_b200 import ops  # noqa: E402


def timed(fn, flush, iters=20):
    fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * iters)]
    for i in range(iters):
        flush.zero_()
        ev[2 * i].record()
        fn()
        ev[2 * i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) * 1e3 for i in range(iters))
    return ts[len(ts) // 2], ts[0]


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, W, C = 32, 64, 64, 180
    x = torch.randn(B, H, W, C, generator=g, device=dev) * 2 + 0.5
    gm = torch.randn(C, generator=g, device=dev) * 0.2 + 1
    bt = torch.randn(C, generator=g, device=dev) * 0.1
    dy = torch.randn(B, H, W, C, generator=g, device=dev)
    dres = torch.randn(B, H, W, C, generator=g, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print({k: os.environ.get(k, "-") for k in ("NSR_LN_V3",)})
    ys, mu, rs = ops.layernorm_fwd(x, gm, bt, sti_out=True, f32_out=False)
    ref = torch.nn.functional.layer_norm(x.double(), (C,), gm.double(), bt.double(), 1e-5)
    err = (ys.to_f32().double() - ref).abs().max().item() / ref.abs().max().item()
    (yf, ys2), _, _ = ops.layernorm_fwd(x, gm, bt, sti_out=True, f32_out=True)
    err32 = (yf.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"fwd  STI-only   median {timed(lambda: ops.layernorm_fwd(x, gm, bt, sti_out=True, f32_out=False), flush)[0]:6.1f} us   "
          f"rel err STI {err:.2e} fp32 {err32:.2e}")
    assert err < 2e-5 and err32 < 2e-6
    dg, db = torch.empty(C, device=dev), torch.empty(C, device=dev)
    xr = x.double().requires_grad_(True)
    gr, br = gm.double().requires_grad_(True), bt.double().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (C,), gr, br, 1e-5).backward(dy.double())
    dx = ops.layernorm_bwd(dy, x, gm, mu, rs, dg, db, dres=dres)
    e1 = ((dx.double() - dres.double()) - xr.grad).abs().max().item() / xr.grad.abs().max().item()
    e2 = (dg.double() - gr.grad).abs().max().item() / gr.grad.abs().max().item()
    e3 = (db.double() - br.grad).abs().max().item() / br.grad.abs().max().item()
    print(f"bwd  fp32 + res median {timed(lambda: ops.layernorm_bwd(dy, x, gm, mu, rs, dg, db, dres=dres), flush)[0]:6.1f} us   "
          f"rel err dx {e1:.2e} dgamma {e2:.2e} dbeta {e3:.2e}")
    print(f"bwd  fp32+STI   median {timed(lambda: ops.layernorm_bwd(dy, x, gm, mu, rs, dg, db, dres=dres, sti_out=True), flush)[0]:6.1f} us")
    rsti = ops.STI.from_f32(dres)
    lean = ops.layernorm_bwd(dy, x, gm, mu, rs, dg, db, dres=rsti, sti_out=True, f32_out=False)
    e4 = ((lean.to_f32().double() - dres.double()) - xr.grad).abs().max().item() / xr.grad.abs().max().item()
    print(f"bwd  STI res->STI median {timed(lambda: ops.layernorm_bwd(dy, x, gm, mu, rs, dg, db, dres=rsti, sti_out=True, f32_out=False), flush)[0]:6.1f} us"
          f"   rel err dx {e4:.2e}   (the form the Swin blocks use)")
    assert e4 < 3e-5
    assert e1 < 1e-5 and e2 < 1e-4 and e3 < 1e-4, (e1, e2, e3)


if __name__ == "__main__":
    main()
