"""`otf` model: `image` whose feed_data synthesises the LQ batch from GT on the device — drop-in for
neosr/models/otf.py (Real-ESRGAN second-order degradation + training-pair pool).

Every stage is one kernel of csrc/otf.cu behind the C ABI.  The host only draws the *plan* — the same
decisions the reference draws (otf.py:111-125,159-178,219-221; transforms.py:93-94) in the same order
from python `random` and a numpy Generator — and uploads the per-sample scalars (sigma / Poisson scale
/ gray flags / JPEG qualities, which the reference draws with device RNG calls) in ONE small copy.
Nothing in feed_data synchronises the host (the reference does 2·B+ syncs: torch.unique loops in the
Poisson noise and the per-sample quality_to_factor loop).
"""
from __future__ import annotations

import random as _random
from typing import Any

import numpy as np
import torch
from torch import Tensor

from .. import ops
from ..registry import MODEL_REGISTRY
from .image import host_default_device, image

MODES = ("area", "bilinear", "bicubic")


def draw_plan(ds: dict, batch: int, ori_h: int, ori_w: int, scale: int, rng: np.random.Generator,
              pyrandom=_random, rng_dev: np.random.Generator | None = None) -> dict:
    """All random decisions of one otf.feed_data call.  The host decisions (python `random` + the numpy
    Generator `rng`) are drawn in the reference's order, so equal seeds give the reference's decisions;
    the per-sample scalars the reference draws with device RNG calls come from `rng_dev`."""
    p: dict[str, Any] = {}
    rd = rng_dev if rng_dev is not None else rng

    def updown(prob_key, range_key):
        t = pyrandom.choices(["up", "down", "keep"], ds.get(prob_key))[0]
        if t == "up":
            return float(rng.uniform(1, ds.get(range_key)[1]))
        if t == "down":
            return float(rng.uniform(ds.get(range_key)[0], 1))
        return 1

    def noise(i, sfx):
        p[f"gauss{i}"] = bool(rng.uniform() < ds.get("gaussian_noise_prob" + sfx))
        lo, hi = ds.get("noise_range" + sfx) if p[f"gauss{i}"] else ds.get("poisson_scale_range" + sfx)
        key = f"sigma{i}" if p[f"gauss{i}"] else f"pscale{i}"
        # per-sample draws (the reference uses torch.rand on the device, degradations.py:654-662,840-848)
        p[key] = (rd.random(batch, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)
        p[f"gray{i}"] = (rd.random(batch, dtype=np.float32) < ds.get("gray_noise_prob" + sfx)).astype(np.float32)

    def jpeg_q(key):
        lo, hi = ds.get(key)
        return (rd.random(batch, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)

    p["scale1"] = updown("resize_prob", "resize_range")
    p["mode1"] = pyrandom.choice(MODES)
    noise(1, "")
    p["jpeg_q1"] = jpeg_q("jpeg_range")
    p["blur2"] = bool(rng.uniform() < ds.get("second_blur_prob"))
    p["scale2"] = updown("resize_prob2", "resize_range2")
    p["mode2"] = pyrandom.choice(MODES)
    noise(2, "2")
    p["sinc_first"] = bool(rng.uniform() < 0.5)
    p["mode3"] = pyrandom.choice(MODES)
    p["jpeg_q2"] = jpeg_q("jpeg_range2")
    ps = ds.get("patch_size")
    h_lq, w_lq = ori_h // scale, ori_w // scale
    if h_lq < ps or w_lq < ps:
        raise ValueError(f"LQ ({h_lq}, {w_lq}) is smaller than patch size ({ps}, {ps}).")
    p["patch_size"] = ps
    p["top"] = pyrandom.randint(0, h_lq - ps)
    p["left"] = pyrandom.randint(0, w_lq - ps)
    p["seed"] = int(rd.integers(0, 2**62))
    return p


@MODEL_REGISTRY.register()
@host_default_device
class otf(image):
    """On The Fly degradations, based on the RealESRGAN pipeline (neosr/models/otf.py:23-35)."""

    def __init__(self, opt: dict[str, Any]) -> None:
        super().__init__(opt)
        ds = dict(opt["datasets"]["train"])
        if opt.get("degradations") is not None:  # train.py:68-70 merges the table into the dataset options
            ds.update(opt["degradations"])
        self._ds = ds
        queue = ds.get("queue_size", 180)
        batch = ds["batch_size"]
        self.queue_size: int = (queue // batch) * batch
        self.patch_size = ds.get("patch_size")
        seed = int(opt.get("manual_seed", 1024) or 1024) + int(opt.get("rank", 0))
        self._rng = np.random.default_rng(seed)
        self._rng_dev = np.random.default_rng([seed, 1])
        self._pyrandom = _random.Random(seed)
        self.queue_ptr = 0
        self.queue_lr = self.queue_gt = None
        self._perm = np.arange(self.queue_size)  # logical pool position -> physical slot

    # -------------------------------------------------------------- pipeline (otf.py:105-257)
    def _per_sample(self, plan: dict) -> dict[str, Tensor]:
        """One pinned upload for every per-sample scalar of the plan."""
        keys = [k for k in ("sigma1", "pscale1", "gray1", "jpeg_q1", "sigma2", "pscale2", "gray2", "jpeg_q2") if k in plan]
        host = torch.from_numpy(np.stack([np.asarray(plan[k], dtype=np.float32) for k in keys])).pin_memory()
        dev = host.to(self.device, non_blocking=True)
        return {k: dev[i] for i, k in enumerate(keys)}

    def run_plan(self, gt: Tensor, kernel1: Tensor, kernel2: Tensor, sinc_kernel: Tensor, plan: dict,
                 fields: dict | None = None):
        """GT batch -> (LQ crop, GT crop) on the device.  `fields` (tests only): caller-provided random
        fields replacing the in-kernel Philox draws (keys z1, zg1, cc1, cg1, z2, zg2, cc2, cg2)."""
        f = fields or {}
        scale = self.opt["scale"]
        dv = self._per_sample(plan)
        ori_h, ori_w = gt.shape[2:]

        def noise(x, i):
            gray = dv[f"gray{i}"]
            any_gray = bool(np.asarray(plan[f"gray{i}"]).sum() > 0)
            seed = plan["seed"] + 7919 * i
            if plan[f"gauss{i}"]:
                return ops.gaussian_noise(x, dv[f"sigma{i}"], gray, any_gray, seed, f.get(f"z{i}"), f.get(f"zg{i}"))
            return ops.poisson_noise(x, dv[f"pscale{i}"], gray, any_gray, seed, f.get(f"cc{i}"), f.get(f"cg{i}"))

        out = ops.filter2d(gt, kernel1)
        if plan["scale1"] != 1:  # scale_factor 1 is the identity in all three modes
            out = ops.resize(out, plan["mode1"], scale_factor=plan["scale1"])
        out = noise(out, 1)
        out = ops.jpeg(out, dv["jpeg_q1"])  # nsr_jpeg clamps its input to [0,1] (otf.py:154,232,239)
        if plan["blur2"]:
            out = ops.filter2d(out, kernel2)
        out = ops.resize(out, plan["mode2"], size=(int(ori_h / scale * plan["scale2"]), int(ori_w / scale * plan["scale2"])))
        out = noise(out, 2)
        final = (ori_h // scale, ori_w // scale)
        if plan["sinc_first"]:
            out = ops.filter2d(ops.resize(out, plan["mode3"], size=final), sinc_kernel)
            out = ops.jpeg(out, dv["jpeg_q2"])
        else:
            out = ops.jpeg(out, dv["jpeg_q2"])
            out = ops.filter2d(ops.resize(out, plan["mode3"], size=final), sinc_kernel)
        ps, top, left = plan["patch_size"], plan["top"], plan["left"]
        lq = ops.crop(out, top, left, ps, ps, quantise=True)  # otf.py:251 fused with the crop
        gt = ops.crop(gt, top * scale, left * scale, ps * scale, ps * scale)
        return lq, gt

    @torch.no_grad()
    def _dequeue_and_enqueue(self, lq: Tensor, gt: Tensor, perm: np.ndarray | None = None):
        """Training-pair pool (otf.py:37-90).  The reference gathers the WHOLE pool through a fresh
        randperm every iteration and swaps the first b entries; the same samples leave and enter when
        only a logical->physical permutation is kept on the host and b slots are swapped in place."""
        b = lq.size(0)
        if self.queue_lr is None:
            assert self.queue_size % b == 0, f"queue size {self.queue_size} should be divisible by batch size {b}"
            self.queue_lr = torch.zeros(self.queue_size, *lq.shape[1:], dtype=torch.float32, device=self.device)
            self.queue_gt = torch.zeros(self.queue_size, *gt.shape[1:], dtype=torch.float32, device=self.device)
        full = self.queue_ptr == self.queue_size
        if full:
            idx = self._rng_dev.permutation(self.queue_size) if perm is None else np.asarray(perm)
            self._perm = self._perm[idx]
            slots_h = self._perm[:b]
        else:
            slots_h = np.arange(self.queue_ptr, self.queue_ptr + b)
            self.queue_ptr += b
        slots = torch.from_numpy(slots_h.astype(np.int32)).pin_memory().to(self.device, non_blocking=True)
        lq_d = ops.pool_swap(self.queue_lr, lq, slots, dequeue=full)
        gt_d = ops.pool_swap(self.queue_gt, gt, slots, dequeue=full)
        return (lq_d, gt_d) if full else (lq, gt)

    @torch.no_grad()
    def feed_data(self, data: dict) -> None:  # otf.py:92-283
        if not self.is_train:
            return super().feed_data(data)
        to = lambda t: t.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()  # noqa: E731
        gt, k1, k2, sk = to(data["gt"]), to(data["kernel1"]), to(data["kernel2"]), to(data["sinc_kernel"])
        plan = draw_plan(self._ds, gt.size(0), gt.size(2), gt.size(3), self.opt["scale"], self._rng, self._pyrandom, self._rng_dev)
        lq, gt = self.run_plan(gt, k1, k2, sk, plan)
        lq, gt = self._dequeue_and_enqueue(lq, gt)
        super().feed_data({"lq": lq, "gt": gt})  # applies apply_augment when configured (otf.py:266-278)
