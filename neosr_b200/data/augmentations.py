"""apply_augment on the device — drop-in for neosr/data/augmentations.py:219-310 (mixup / cutmix / resizemix / cutblur
on the LQ batch up-sampled to GT resolution).  The host draws the plan (same draws, same order, from python `random`
and the numpy Generator, as the reference; the batch permutations the reference draws with torch.randperm come from
`rng_dev`); each operation is one launch of `nsr_resize_aa` / `nsr_batch_mix`."""
from __future__ import annotations

import random as _random

import numpy as np
import torch
from torch import Tensor

from .. import ops

ORDER = ("cutmix", "mixup", "resizemix", "cutblur")  # application order in the multi-augmentation branch (278-285)


def _bbox(size2: int, size3: int, cut_w: int, cut_h: int, rng):
    """rand_bbox of cutmix / resizemix / cutblur (41-54, 82-95, 143-156): W = size[2], H = size[3]."""
    cx, cy = int(rng.integers(size2)), int(rng.integers(size3))
    return (int(np.clip(cx - cut_w // 2, 0, size2)), int(np.clip(cy - cut_h // 2, 0, size3)),
            int(np.clip(cx + cut_w // 2, 0, size2)), int(np.clip(cy + cut_h // 2, 0, size3)))


def draw_augment_plan(batch: int, h: int, w: int, scale: int, augs, prob, rng: np.random.Generator, pyrandom=_random,
                      rng_dev: np.random.Generator | None = None, multi_prob: float = 0.3) -> dict:
    if len(augs) != len(prob):
        raise ValueError("Length of 'augmentation' and aug_prob don't match!")
    if batch == 1:
        raise ValueError("Augmentations need batch >1 to work.")
    rd = rng_dev if rng_dev is not None else rng
    plan: dict = {"up_mode": pyrandom.choice(["bilinear", "bicubic"]) if scale > 1 else None, "ops": []}
    if rng.random() < multi_prob:
        num = int(rng.integers(2, len(augs))) if len(augs) > 2 else len(augs)
        weighted = list(zip(augs, prob, strict=False))
        chosen = []
        for _ in range(num):
            c = pyrandom.choices(weighted, k=1)  # uniform over what is left: the reference passes no weights here
            chosen.append(c[0][0])
            weighted.remove(c[0])
        names = [n for n in ORDER if n in chosen]
    else:
        a = augs[pyrandom.choices(range(len(augs)), weights=prob)[0]]
        names = [n for n in ORDER if n in a][:1]
    for n in names:
        if n == "mixup":
            op = {"op": n, "lam": float(rng.uniform(0.4, 0.6)), "perm": rd.permutation(batch)}
        elif n == "cutmix":
            lam = float(rng.uniform(0, 0.9))
            perm = rd.permutation(batch)
            cut = np.sqrt(1.0 - lam)
            op = {"op": n, "perm": perm, "box": _bbox(h, w, int(h * cut), int(w * cut), rng)}
        elif n == "resizemix":
            perm = rd.permutation(batch)
            tao = float(rng.uniform(0.2, 0.9))
            op = {"op": n, "perm": perm, "box": _bbox(h, w, int(h * tao), int(w * tao), rng)}
        else:
            lam = float(rng.uniform(0.2, 0.7))
            op = {"op": n, "box": _bbox(h, w, int(h * lam), int(w * lam), rng)}
        plan["ops"].append(op)
    return plan


def run_augment_plan(gt: Tensor, lq: Tensor, scale: int, plan: dict):
    dev = gt.device
    if scale > 1:
        lq = ops.resize_aa(lq, plan["up_mode"], scale_factor=scale)
    if gt.shape != lq.shape and plan["ops"]:
        raise ValueError("img_gt and img_lq have to be the same resolution.")
    for op in plan["ops"]:
        perm = None
        if "perm" in op:
            perm = torch.from_numpy(np.asarray(op["perm"]).astype(np.int32)).pin_memory().to(dev, non_blocking=True)
        if op["op"] == "mixup":
            gt, lq = ops.batch_mix(gt, gt, perm, "mixup", op["lam"]), ops.batch_mix(lq, gt, perm, "mixup", op["lam"])
        elif op["op"] == "cutmix":
            x1, y1, x2, y2 = op["box"]  # [:, :, bbx1:bbx2, bby1:bby2] (58-59): "x" runs along dim 2
            gt, lq = ops.batch_mix(gt, gt, perm, "box", box=(x1, x2, y1, y2)), ops.batch_mix(lq, lq, perm, "box", box=(x1, x2, y1, y2))
        elif op["op"] == "resizemix":
            x1, y1, x2, y2 = op["box"]  # pasted at [:, :, bby1:bby2, bbx1:bbx2] (121-122)
            if y2 > y1 and x2 > x1:
                gt2, lq2 = gt.clone(), lq.clone()
                ops.resize_aa(gt, "bicubic", size=(y2 - y1, x2 - x1), perm=perm, dst=gt2, top=y1, left=x1)
                ops.resize_aa(lq, "bicubic", size=(y2 - y1, x2 - x1), perm=perm, dst=lq2, top=y1, left=x1)
                gt, lq = gt2, lq2
        else:  # cutblur: img_lq[box] = img_gt[box] (164)
            x1, y1, x2, y2 = op["box"]
            lq = ops.batch_mix(lq, gt, None, "box", box=(x1, x2, y1, y2))
    if scale > 1:
        lq = ops.resize_aa(lq, "bicubic", scale_factor=1 / scale)
    return gt, lq
