import sys, torch
sys.path.insert(0, ".")
import bench
from neosr_b200.models import build_model
cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.cuda.set_device(0)
opt = bench.make_opt(B, False, 0, 1, cfg)
model = build_model(opt)
pool = bench.synth_otf_batches(4, B, 1024, 128 if cfg == "c4" else 192)
for it in range(10):
    model.feed_data(pool[it % 4])
    lq, gt = model.lq, model.gt
    print(it, "lq", float(lq.min()), float(lq.max()), bool(torch.isnan(lq).any()), "gt", float(gt.min()), float(gt.max()))
    try:
        model.optimize_parameters(it)
        print("   ", {k: round(v, 5) for k, v in model.get_current_log().items()})
    except ValueError as e:
        logs = model._pending_logs
        print("   NaN:", {k: float(v.flatten()[0]) for k, v in logs.items()})
        out = model.output
        print("   out", float(out.min()), float(out.max()), bool(torch.isnan(out).any()))
        break
