"""One EAGER training step of a bench configuration between cudaProfilerStart/Stop, for an ncu pass over EVERY kernel
of the step (run under `ncu --profile-from-start off --metrics ... --csv`):
    python tools/one_step.py --config c3|c2|c4|c5 [--batch B]
"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--batch", type=int, default=0)
    a = ap.parse_args()
    from neosr_b200.models import build_model
    B = a.batch or bench.DEFAULT_BATCH[a.config]
    opt = bench.make_opt(B, False, 0, 1, a.config)
    opt["cuda_graph"] = False
    model = build_model(opt)
    if a.config in ("c4", "c5"):
        pool = bench.synth_otf_batches(2, B, seed=1024, hr=128 if a.config == "c4" else 192)
    else:
        pool = bench.synth_batches(2, B, seed=1024)
    for i in range(2):
        model.feed_data(pool[i % 2])
        model.optimize_parameters(i + 1)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model.feed_data(pool[0])
    model.optimize_parameters(3)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("loss", model.get_current_log())


if __name__ == "__main__":
    main()
