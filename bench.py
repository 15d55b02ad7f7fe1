#!/usr/bin/env python
"""Headline benchmark: LR-crops/s through one full training step (feed_data +
optimize_parameters) of SwinIR-medium x4, 64->256 RGB crops, L1 + VGG19-perceptual,
adan_sf + EMA, batch 32 per GPU (BASELINE.json configs[2], "C3").

    python bench.py --gpus N --steps K --warmup W            # our arm (B200 kernels, C ABI)
    python bench.py --impl reference --gpus N ...             # reference arm: the reference's own step
                                                              # (baseline/_ref) on the host CPU cores
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# algorithmic work per LR crop for C3 (SURVEY.md §8d / BASELINE.md §3): G fwd+bwd + VGG fwd(x)+fwd(gt)+dgrad(x)
GFLOP_PER_CROP_C3 = 321.30 + 152.88
# C2 G 440.56 + VGG 152.88 + D 103.68 + 2*155.52; C4 (hat_l, HR 128) 446.4; C5 (realplksr, HR 192) 421.3
GFLOP_PER_CROP = {"c3": GFLOP_PER_CROP_C3, "c2": 1008.2, "c4": 446.4, "c5": 421.3}
DEFAULT_BATCH = {"c3": 32, "c2": 16, "c4": 8, "c5": 64}
METRIC = "LR-crops/sec (SwinIR-M 4x, 64->256, full training step)"
METRICS = {"c3": METRIC, "c2": "LR-crops/sec (ESRGAN 4x GAN step, 64->256)",
           "c4": "LR-crops/sec (HAT-L 4x otf step, HR 128 -> LQ 32)", "c5": "LR-crops/sec (RealPLKSR 4x otf step, HR 192 -> LQ 48)"}


def peaks() -> dict:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


WORKLOADS = {
    "c3": "C3: swinir_medium x4, 64x64->256x256 RGB, L1(1.0)+vgg19 perceptual(0.5, chc), adan_sf + EMA 0.999, "
          "grad-clip 1.0, drop_path 0",
    "c2": "C2: esrgan (23 RRDB) x4, 64x64->256x256 RGB, L1(1.0)+vgg19 perceptual(0.5, chc)+GAN(bce 0.1, unet "
          "discriminator with spectral norm), adan_sf on G and D + EMA 0.999, grad-clip 1.0",
    "c4": "C4: hat_l x4, `otf` model: on-the-fly degradation of 128x128 HR crops (train_hat_otf.toml [degradations]) -> "
          "32x32 LQ, pool 176, mssim(1.0)+consistency(1.0)+vgg19 perceptual(0.5)+GAN(bce 0.3, unet), adan_sf + EMA",
    "c5": "C5: realplksr x4, `otf` model: on-the-fly degradation of 192x192 HR crops -> 48x48 LQ, pool 128, "
          "mssim(1.0)+consistency(1.0)+vgg19 perceptual(0.5)+GAN(bce 0.2, unet), AdamW lr 1e-4 (template: 5e-4; see make_opt) "
          "+ EMA 0.999, drop_path 0",
}


def make_opt(batch: int, dist: bool, rank: int, world: int, config: str = "c3") -> dict:
    if config == "c2":
        o = make_opt(batch, dist, rank, world, "c3")
        o["name"] = "bench_c2"
        o["network_g"] = {"type": "esrgan", "num_block": 23, "num_feat": 64, "num_grow_ch": 32}
        o["network_d"] = {"type": "unet", "num_feat": 64}
        o["train"]["optim_d"] = dict(o["train"]["optim_g"])
        o["train"]["gan_opt"] = {"type": "gan_loss", "gan_type": "bce", "loss_weight": 0.1}
        return o
    if config in ("c4", "c5"):
        from neosr_b200.data.degradations import TEMPLATE_DEGRADATIONS as DEGRADATIONS
        o = make_opt(batch, dist, rank, world, "c3")
        ps = 32 if config == "c4" else 48
        o.update(name=f"bench_{config}", model_type="otf", manual_seed=1024)
        o["datasets"]["train"] = dict(DEGRADATIONS, patch_size=ps, batch_size=batch, queue_size=180)
        o["network_d"] = {"type": "unet", "num_feat": 64}
        tr = o["train"]
        del tr["pixel_opt"]
        tr["mssim_opt"] = {"type": "mssim_loss", "loss_weight": 1.0}
        tr["consistency_opt"] = {"type": "consistency_loss", "loss_weight": 1.0}
        tr["gan_opt"] = {"type": "gan_loss", "gan_type": "bce", "loss_weight": 0.3 if config == "c4" else 0.2}
        if config == "c4":
            o["network_g"] = {"type": "hat_l", "drop_path_rate": 0.0, "upscale": 4}
            tr["optim_d"] = dict(tr["optim_g"])
        else:
            o["network_g"] = {"type": "realplksr", "upscaling_factor": 4}
            # lr 1e-4: at 5e-4 a random-init generator occasionally drives an MS-SSIM scale's cs mean negative within
            # the bench's ~25 steps and the loss goes NaN (as it would in the reference, ssim_loss.py:131-144)
            tr["optim_g"] = {"type": "AdamW", "lr": 1e-4, "betas": [0.9, 0.99], "weight_decay": 0.01}
            tr["optim_d"] = dict(tr["optim_g"])
        return o
    return {"name": "bench_c3", "model_type": "image", "scale": 4, "is_train": True, "dist": dist, "rank": rank,
            "world_size": world, "num_gpu": world,
            "network_g": {"type": "swinir_medium", "drop_path_rate": 0.0, "upscale": 4},
            "datasets": {"train": {"patch_size": 64, "batch_size": batch}},
            "train": {"ema": 0.999,
                      "optim_g": {"type": "adan_sf", "lr": 1e-3, "betas": [0.98, 0.92, 0.987], "weight_decay": 0.02,
                                  "schedule_free": True, "warmup_steps": 1600},
                      "pixel_opt": {"type": "L1Loss", "loss_weight": 1.0},
                      "perceptual_opt": {"type": "vgg_perceptual_loss", "loss_weight": 0.5, "criterion": "chc",
                                         "allow_random_init": True}},
            "path": {}}


def synth_batches(n: int, batch: int, seed: int, lq_size: int = 64, scale: int = 4):
    """SURVEY.md §8d synthetic inputs: gt = rand quantised to 8 bit; lq = antialiased bicubic
    downsample, clamped, quantised.  Returned in pinned host memory."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = []
    for _ in range(n):
        gt = torch.rand(batch, 3, lq_size * scale, lq_size * scale, generator=g)
        gt = torch.round(gt * 255) / 255
        lq = F.interpolate(gt, scale_factor=1 / scale, mode="bicubic", antialias=True).clamp(0, 1)
        lq = torch.round(lq * 255) / 255
        out.append({"lq": lq.contiguous().pin_memory() if torch.cuda.is_available() else lq,
                    "gt": gt.contiguous().pin_memory() if torch.cuda.is_available() else gt})
    return out


def synth_otf_batches(n: int, batch: int, seed: int, hr: int):
    """OTF inputs: STRUCTURED synthetic GT on 8-bit levels (low-pass noise + edges: white noise blurs to
    a constant and drives MS-SSIM's cs mean negative, whose fractional power is NaN in the reference too) + the three
    blur kernels per sample from the host synthesis (neosr_b200/data/degradations.py)."""
    import random

    import numpy as np
    import torch

    from neosr_b200.data.degradations import TEMPLATE_DEGRADATIONS as DEGRADATIONS
    from neosr_b200.data.degradations import synth_kernels
    from neosr_b200.data.synthetic import structured_gt
    rng, pr = np.random.default_rng(seed), random.Random(seed)
    out = []
    for i in range(n):
        gt = structured_gt(seed * 1009 + i, batch, hr, hr)
        ks = [synth_kernels(DEGRADATIONS, rng, pr) for _ in range(batch)]
        d = {"gt": gt, "kernel1": torch.from_numpy(np.stack([k[0] for k in ks])),
             "kernel2": torch.from_numpy(np.stack([k[1] for k in ks])), "sinc_kernel": torch.from_numpy(np.stack([k[2] for k in ks]))}
        out.append({k: (v.contiguous().pin_memory() if torch.cuda.is_available() else v) for k, v in d.items()})
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                r = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                    str(self.index)], capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([c.strip() for c in r.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self) -> dict:
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------- reference arm
def _reference_available() -> bool:
    from oracle import ref_shim
    return ref_shim.available()


def _reference_model(device: str, tf32: bool = False):
    """The UNMODIFIED reference (baseline/_ref, or /root/reference in the build container): its own `swinir_medium`,
    `L1Loss`, `vgg_perceptual_loss` (seeded-random VGG19: no pretrained weights offline, as on our arm), `adan_sf` and
    EMA, driven through its REAL `image.feed_data` / `image.optimize_parameters` (neosr/models/image.py:374-391,
    427-662) on an `object.__new__(image)` instance (the constructors hard-wire cuda + DDP)."""
    import torch

    from oracle import ref_shim
    ref_shim.activate(4)
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    if tf32:  # what `fast_matmul = true` does (train.py:168-173)
        torch.set_float32_matmul_precision("medium")
    else:
        torch.set_float32_matmul_precision("highest")
    torch.manual_seed(0)
    net = ref_shim.build_network({"type": "swinir_medium", "drop_path_rate": 0.0})
    from neosr.losses.basic_loss import L1Loss
    percep = ref_shim.build_vgg_perceptual(None, loss_weight=0.5, criterion="chc")
    return ref_shim.make_image_model(net, cri_pix=L1Loss(loss_weight=1.0), cri_perceptual=percep, ema=0.999, scale=4,
                                     optim_kw=dict(lr=1e-3, betas=(0.98, 0.92, 0.987), weight_decay=0.02,
                                                   schedule_free=True, warmup_steps=1600), device=device)


def reference_cpu_run(batch: int, steps: int, warmup: int, pool: int = 4) -> dict:
    """The reference's own CPU step on all host threads, bounded sample = `batch` crops per step (crops/s is
    batch-normalised; BASELINE.md section 4 / SURVEY.md section 8d: B = 4)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = _reference_model("cpu")
    data = synth_batches(pool, batch, seed=1024)
    for i in range(warmup):
        m.feed_data(data[i % pool])
        m.optimize_parameters(i + 1)
    t0 = time.perf_counter()
    for i in range(steps):
        m.feed_data(data[i % pool])
        m.optimize_parameters(warmup + i + 1)
    mean = (time.perf_counter() - t0) / steps
    return {"value": batch / mean, "unit": "crops/s", "cores": cores, "kind": "reference",
            "sample": f"{steps} timed + {warmup} warm-up step(s) of the reference's real image.feed_data + "
                      f"image.optimize_parameters (swinir_medium, L1 + vgg19 perceptual, adan_sf + EMA) at batch {batch} "
                      f"(crops/s is batch-normalised), torch CPU fp32, {cores} threads", "ms_per_step": mean * 1e3,
            "steps": steps, "loss": float(m.log_dict.get("l_g_total", float("nan")))}


def reference_cuda_run(batch: int, steps: int, warmup: int, tf32: bool, pool: int = 4) -> dict:
    """Informational: the reference's eager CUDA step on the same B200 (what a user of the reference sees today),
    fp32 (TF32 off) or with `fast_matmul` (TF32 on).  Device-timed with CUDA events, inputs fed from pinned memory."""
    import torch
    data = synth_batches(pool, batch, seed=1024)
    torch.set_default_device("cuda")  # what the reference's entry point does before it builds anything (train.py:165)
    try:
        return _reference_cuda_steps(data, batch, steps, warmup, tf32, pool)
    finally:
        torch.set_default_device("cpu")


def _reference_cuda_steps(data, batch: int, steps: int, warmup: int, tf32: bool, pool: int) -> dict:
    import torch
    m = _reference_model("cuda", tf32=tf32)
    for i in range(warmup):
        m.feed_data(data[i % pool])
        m.optimize_parameters(i + 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        m.feed_data(data[i % pool])
        m.optimize_parameters(warmup + i + 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"value": batch / (ms * 1e-3), "unit": "crops/s", "ms_per_step": ms, "batch": batch, "steps": steps,
           "warmup": warmup, "tf32": tf32, "loss": float(m.log_dict.get("l_g_total", float("nan"))),
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    del m
    torch.cuda.empty_cache()
    return out


def cpu_oracle_run(batch: int, steps: int, warmup: int, budget_s: float) -> dict:
    """Fallback when no reference tree is reachable: the oracle's CPU restatement of the same step."""
    import torch

    from oracle import losses as OL
    from oracle.step import make_swinir_trainer
    from oracle.swinir import swinir_medium_config, swinir_param_shapes, synth_params
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = swinir_medium_config(4)
    p = synth_params(swinir_param_shapes(cfg), seed=0)
    vgg_p = synth_params(OL.vgg19_conv_shapes(), seed=5)
    tr = make_swinir_trainer(p, cfg, pixel_weight=1.0, percep_weight=0.5, vgg_params=vgg_p, ema=0.999,
                             optim=dict(lr=1e-3, betas=(0.98, 0.92, 0.987), weight_decay=0.02, schedule_free=True,
                                        warmup_steps=1600))
    data = synth_batches(2, batch, seed=1024)
    t_start = time.perf_counter()
    for i in range(warmup):
        tr.feed_data(data[i % 2])
        tr.optimize_parameters(i)
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        tr.feed_data(data[i % 2])
        tr.optimize_parameters(warmup + i)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    mean = sum(times) / len(times)
    return {"value": batch / mean, "unit": "crops/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} timed + {warmup} warm-up step(s) of the full C3 step at batch {batch} (crops/s is "
                      f"batch-normalised), torch CPU fp32, {cores} threads", "ms_per_step": mean * 1e3,
            "steps": len(times)}


def cpu_baseline_run(batch: int, steps: int, warmup: int, budget_s: float) -> dict:
    if _reference_available():
        return reference_cpu_run(batch, steps, warmup)
    return cpu_oracle_run(1, steps, warmup, budget_s)


def run_reference(args) -> None:
    """`--impl reference`: the reference's own CPU implementation of the step on the box's host cores (rank 0 only),
    same metric / unit / workload string as our arm; plus, when a GPU is visible, the reference's eager CUDA step
    at the full batch (fp32 and fast_matmul/TF32) as informational keys."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b = args.batch or 4
    r = cpu_baseline_run(b, args.steps, args.warmup, budget_s=240.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "crops/s", "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS["c3"], "batch_per_step": b,
                       "arm": "the reference's own CPU step (neosr image.optimize_parameters), bounded sample (see "
                              "cpu_baseline.sample)" if r["kind"] == "reference" else
                              "CPU oracle port of the same step (no reference tree reachable)"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if r["kind"] == "reference" and not args.no_ref_cuda:
        try:
            import torch
            if torch.cuda.is_available():
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
                for key, tf32 in (("ref_cuda_fp32", False), ("ref_cuda_tf32", True)):
                    line[key] = reference_cuda_run(DEFAULT_BATCH["c3"], min(args.steps, 8), min(args.warmup, 3), tf32)
        except Exception as e:  # informational leg: never lose the line
            import traceback
            line["ref_cuda_error"] = (f"{type(e).__name__}: {e}"[:300] + " | " +
                                      " <- ".join(f"{f.name}:{f.lineno}" for f in traceback.extract_tb(e.__traceback__)[-6:]))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- our arm
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    from neosr_b200 import ops
    from neosr_b200.models import build_model
    B = args.batch if args.batch else DEFAULT_BATCH[args.config]
    opt = make_opt(B, world > 1, rank, world, args.config)
    if args.amp:  # opt-in mixed precision (`use_amp` + `bfloat16`): NOT the headline metric, which is fp32 semantics
        opt["use_amp"], opt["bfloat16"] = True, True
    opt["cuda_graph"] = not args.no_graph
    model = build_model(opt)
    n_pool = args.pool or (64 if args.config == "c3" else 8)
    if args.config in ("c4", "c5"):
        pool = synth_otf_batches(n_pool, B, seed=1024 + rank, hr=128 if args.config == "c4" else 192)
    else:
        pool = synth_batches(n_pool, B, seed=1024 + rank)
    dev_pool = [{k: v.cuda(non_blocking=True) for k, v in b.items()} for b in pool]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(loop, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            loop(i)
        e1.record()
        host_ms[0] = (time.perf_counter() - t0) * 1e3 / steps  # host time to ENQUEUE a step
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    it = [0]

    def step_resident(i):
        model.feed_data(dev_pool[i % len(dev_pool)])
        model.optimize_parameters(it[0])
        it[0] += 1

    losses = []

    def step_e2e(i):
        model.feed_data(pool[i % len(pool)])       # pinned host -> device inside the timed region
        model.optimize_parameters(it[0])
        it[0] += 1
        losses.append(model.get_current_log()["l_g_total"])  # device -> host read of the step's loss

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.LAUNCHES
    ms = timed(step_resident, args.steps)
    host_enqueue_ms = host_ms[0]
    launches = ops.LAUNCHES - l0
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # one extra EAGER step with per-launch CUDA events: which kernel dominates, and its roofline
    graph_mode, model._graph_mode = model._graph_mode, False
    ops.PROFILE = []
    step_resident(0)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    model._graph_mode = graph_mode
    agg: dict = {}
    for name, key, flops, nbytes, a, b in prof:
        r = agg.setdefault((name, key), {"ms": 0.0, "n": 0, "flops": flops, "bytes": nbytes})
        r["ms"] += a.elapsed_time(b)
        r["n"] += 1
    step_ms_prof = sum(r["ms"] for r in agg.values())
    fam: dict = {}
    for (name, key), r in agg.items():
        f = fam.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
        f["ms"] += r["ms"]; f["n"] += r["n"]; f["flops"] += r["flops"] * r["n"]; f["bytes"] += r["bytes"] * r["n"]
    pk = peaks()
    # kernel families = one CUDA kernel each: the tcgen05 implicit GEMM serves Linear/Conv fprop AND dgrad
    # (fp32-A and STI-A variants), wgrad is its own kernel, <=4-channel convs run in conv_small.cu
    def family(name, key):
        if name.startswith("conv_"):
            small = min(key[1], key[2]) < 16
            if small:
                return "conv_small (3-channel image-side convs)"
            if "wgrad" in name:
                return "igemm_wgrad_tc (tcgen05)"
            # two instantiations of the kernel template, two different kernels in the binary: <BN, STI=true> takes
            # bulk-copied split-tile operands (the 1x1 contractions of the transformer body), <BN, STI=false> gathers the
            # im2col tile with producer warps (3x3 convs)
            return ("igemm_fprop_tc<STI> (tcgen05, 1x1 fprop+dgrad, bulk-copied operands)" if name.endswith("_sti")
                    else "igemm_fprop_tc (tcgen05, 3x3 fprop+dgrad, producer warps)")
        return name
    fams: dict = {}
    for (name, key), r in agg.items():
        f = fams.setdefault(family(name, key), {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
        f["ms"] += r["ms"]; f["n"] += r["n"]; f["flops"] += r["flops"] * r["n"]; f["bytes"] += r["bytes"] * r["n"]
    dname, drec = max(fams.items(), key=lambda kv: kv[1]["ms"])
    avg_ms = drec["ms"] / drec["n"]
    # The binding roof of a kernel family is the larger of its two minimal times over one step's launches:
    #   t_hbm = algorithmic bytes / measured HBM bandwidth,  t_tensor = MMA passes x algorithmic FLOPs / measured bf16 peak
    # (3 passes: hi*hi + hi*lo + lo*hi of the split-bf16 scheme that fp32 parity needs).  frac = that time / measured time.
    t_meas = drec["ms"] * 1e-3
    t_hbm = drec["bytes"] / (pk["hbm_gbs"] * 1e9)
    t_tensor = 3.0 * drec["flops"] / (pk["bf16_sustained"] * 1e12)
    if t_tensor >= t_hbm and drec["flops"] > 0:
        ach = drec["flops"] / t_meas / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_sustained"], "traffic": None, "mma_passes": 3,
                "tensor_pipe_frac": 3.0 * ach / pk["bf16_sustained"],
                "note": "achieved = algorithmic (fp32-equivalent) FLOPs / device time over all launches of this kernel in one "
                        "step; each product is issued as 3 bf16 MMA passes (hi*hi+hi*lo+lo*hi), so the tensor pipe does 3x this "
                        "(tensor_pipe_frac); peak = sustained bf16 (kernel timed inside a long step)"}
    else:
        ach = drec["bytes"] / t_meas / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": None,
                "note": "achieved = algorithmic bytes (4 B x pixels x (cin + cout) per contraction; SURVEY.md §8d) / device time "
                        "over all launches of this kernel in one step; at K = 180..540 the Swin contractions sit below the "
                        "ridge (35 FLOP/B vs 73 FLOP/B for 3-pass bf16), so HBM is the binding roof"}
    roof.update({"hbm_time_frac": t_hbm / t_meas, "tensor_time_frac": t_tensor / t_meas})
    try:  # measured DRAM traffic per launch of the dominant kernel, from the committed ncu --set full capture
        tr = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(dname)
        if tr:
            roof["traffic"] = tr["bytes_per_launch"]
            roof["traffic_source"] = tr["source"]
            roof["algorithmic_bytes_per_launch"] = drec["bytes"] / drec["n"]
    except (OSError, ValueError):
        pass
    roof.update({"kernel": dname, "avg_launch_ms": avg_ms, "launches_per_step": drec["n"],
                 "share_of_step": drec["ms"] / step_ms_prof, "peak_source": pk["source"],
                 "engine": os.environ.get("NSR_ENGINE", "auto"),
                 "top_kernels": [{"kernel": k, "ms_per_step": round(v["ms"], 3), "share": round(v["ms"] / step_ms_prof, 4),
                                  "launches": v["n"]} for k, v in sorted(fams.items(), key=lambda kv: -kv[1]["ms"])[:8]]})

    if rank == 0:
        crops = B * world * args.steps
        value = crops / (ms * 1e-3)
        step_tflops = value / world * GFLOP_PER_CROP[args.config] / 1e3
        out_dir = ROOT / "gpurun_out"
        try:
            out_dir.mkdir(exist_ok=True)
            table = sorted(({"kernel": n, "shape": list(k), **r} for (n, k), r in agg.items()), key=lambda r: -r["ms"])
            (out_dir / f"kernel_table_n{world}.json").write_text(json.dumps(
                {"step_ms_sum_of_kernels": step_ms_prof, "families": fam, "kernels": table}, indent=1))
        except OSError:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline and args.config == "c3":
            r = cpu_baseline_run(4, 3, 1, budget_s=60.0)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        bytes_in = sum(v.numel() * 4 for v in pool[0].values())
        line = {"metric": METRICS[args.config], "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16 products, f32 accumulate / storage (use_amp)" if args.amp else "f32",
                "data": "synthetic (seeded rand, 8-bit quantised; VGG19 weights "
                                                               "seeded-random: no pretrained weights offline)",
                "config": {"workload": WORKLOADS[args.config] + (" [use_amp + bfloat16: single bf16 pass]" if args.amp else ""),
                           "batch_per_gpu": B,
                           "global_batch": B * world, "parallelism": f"dp{world}",
                           "l2": "per-step working set (saved activations, GBs) >> 126 MB L2; pool of "
                                 f"{len(pool)} distinct batches cycled"},
                "clocks": sampler.summary(),
                "e2e": {"value": crops / (ms_e2e * 1e-3), "unit": "crops/s", "h2d_bytes_per_step": bytes_in,
                        "d2h_bytes_per_step": 4 * 3, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
                "cuda_graph": bool(graph_mode and model._graphs is not None),
                "roofline": roof,
                "step_roofline": {"bound": "tensor", "achieved": step_tflops, "unit": "TFLOP/s",
                                  "peak": pk["bf16_sustained"], "frac": step_tflops / pk["bf16_sustained"],
                                  "gflop_per_crop": GFLOP_PER_CROP[args.config],
                                  "note": "whole step vs sustained bf16 peak; the fp32-parity engines are exact-fp32 "
                                          "SIMT or 3xBF16-split tcgen05 (ceiling = peak/3)"},
                "cpu_baseline": cpu, "final_loss": losses[-1] if losses else None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="LR crops per GPU per step (default: C3 32, C2 16)")
    ap.add_argument("--config", default="c3", choices=["c3", "c2", "c4", "c5"],
                    help="c3 = BASELINE.json's headline workload (what the driver runs); c2 = the GAN configuration; "
                         "c4 / c5 = the otf configurations (hat_l / realplksr, per-GPU shapes of BASELINE.json configs[3], [4])")
    ap.add_argument("--pool", type=int, default=0, help="distinct synthetic batches cycled (default: 64 for c3 as SURVEY.md section 8d states, 8 for the others)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="reference arm: skip the informational eager-CUDA legs")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly (no CUDA-graph replay)")
    ap.add_argument("--amp", action="store_true", help="opt-in mixed precision (use_amp + bfloat16); informational, not the headline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
