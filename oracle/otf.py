"""CPU oracle for the on-the-fly degradation pipeline (`otf.feed_data`).  TEST INFRASTRUCTURE ONLY.

PyTorch-CPU restatement of neosr/models/otf.py:92-283 and the helpers it calls, with every random
decision passed in as a *plan* (a plain dict, see `neosr_b200.models.otf.draw_plan`) and every random
field (standard-normal / Poisson draws) either passed in or drawn here and returned, so the CUDA path
can be replayed on the very same decisions and fields.

Pinned against the live reference functions (`filter2D`, `DiffJPEG`, `generate_gaussian_noise_pt`,
`generate_poisson_noise_pt`, `paired_random_crop`, `_dequeue_and_enqueue`) by
tests/test_oracle_vs_reference.py and, through fixtures written by oracle/make_golden.py, by
tests/test_oracle_golden.py.  `F.interpolate` is torch itself (third-party arithmetic).
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import Tensor
from torch.nn import functional as F

# standard JPEG tables; the reference stores them transposed (diffjpeg.py:16-38)
_Y = np.array([[16, 11, 10, 16, 24, 40, 51, 61], [12, 12, 14, 19, 26, 58, 60, 55], [14, 13, 16, 24, 40, 57, 69, 56],
               [14, 17, 22, 29, 51, 87, 80, 62], [18, 22, 37, 56, 68, 109, 103, 77], [24, 35, 55, 64, 81, 104, 113, 92],
               [49, 64, 78, 87, 103, 121, 120, 101], [72, 92, 95, 98, 112, 100, 103, 99]], dtype=np.float32).T
_C = np.full((8, 8), 99, dtype=np.float32)
_C[:4, :4] = np.array([[17, 18, 24, 47], [18, 21, 26, 66], [24, 26, 56, 99], [47, 66, 99, 99]], dtype=np.float32).T


def filter2d(img: Tensor, kernel: Tensor) -> Tensor:
    """diffjpeg.py:558-584."""
    k = kernel.size(-1)
    if k % 2 != 1:
        raise ValueError("Wrong kernel size")
    b, c, h, w = img.shape
    p = F.pad(img, (k // 2,) * 4, mode="reflect")
    if kernel.size(0) == 1:
        return F.conv2d(p.reshape(b * c, 1, *p.shape[-2:]), kernel.view(1, 1, k, k)).view(b, c, h, w)
    wgt = kernel.view(b, 1, k, k).repeat(1, c, 1, 1).view(b * c, 1, k, k)
    return F.conv2d(p.reshape(1, b * c, *p.shape[-2:]), wgt, groups=b * c).view(b, c, h, w)


def quality_to_factor(q: Tensor) -> Tensor:
    """diffjpeg.py:48-61, vectorised (the reference loops over samples, 542-543)."""
    return torch.where(q < 50, 5000.0 / q, 200.0 - q * 2) / 100.0


def _dct_basis(dtype):
    x = torch.arange(8, dtype=torch.float64)
    c = torch.cos((2 * x[:, None] + 1) * x[None, :] * math.pi / 16)  # [x, u]
    return c.to(dtype)


def jpeg(x: Tensor, quality: Tensor) -> Tensor:
    """DiffJPEG(differentiable=False) (diffjpeg.py:254-291,461-508,531-555), dtype-generic (fp32 or fp64)."""
    dt = x.dtype
    factor = quality_to_factor(quality.to(dt))
    b, _, h, w = x.shape
    hp, wp = (16 - h % 16) % 16, (16 - w % 16) % 16
    x = F.pad(x, (0, wp, 0, hp)) * 255
    H, W = h + hp, w + wp
    m = torch.tensor([[0.299, 0.587, 0.114], [-0.168736, -0.331264, 0.5], [0.5, -0.418688, -0.081312]],
                     dtype=torch.float32).to(dt)
    ycc = torch.einsum("bchw,kc->bkhw", x, m) + torch.tensor([0.0, 128.0, 128.0], dtype=dt).view(1, 3, 1, 1)
    comps = [ycc[:, 0], F.avg_pool2d(ycc[:, 1:2], 2)[:, 0], F.avg_pool2d(ycc[:, 2:3], 2)[:, 0]]
    cb = _dct_basis(dt)
    alpha = torch.tensor([1 / math.sqrt(2)] + [1.0] * 7, dtype=torch.float64)
    a2 = torch.outer(alpha, alpha).to(dt)
    outs = []
    for i, comp in enumerate(comps):
        hh, ww = comp.shape[1:]
        blk = comp.view(b, hh // 8, 8, ww // 8, 8).permute(0, 1, 3, 2, 4) - 128  # [b, by, bx, x, y]
        coef = torch.einsum("bijxy,xu,yv->bijuv", blk, cb, cb) * (a2 * 0.25)
        tab = torch.from_numpy(_Y if i == 0 else _C).to(dt).view(1, 1, 1, 8, 8) * factor.view(b, 1, 1, 1, 1)
        q = torch.round(coef / tab) * tab
        rec = 0.25 * torch.einsum("bijxy,ux,vy->bijuv", q * a2, cb, cb) + 128
        outs.append(rec.permute(0, 1, 3, 2, 4).reshape(b, hh, ww))
    y, cbp, crp = outs
    up = lambda t: t.repeat_interleave(2, 1).repeat_interleave(2, 2)  # noqa: E731
    ycc = torch.stack([y, up(cbp) - 128, up(crp) - 128], 1)
    mi = torch.tensor([[1.0, 0.0, 1.402], [1, -0.344136, -0.714136], [1, 1.772, 0]], dtype=torch.float32).to(dt)
    rgb = torch.einsum("bchw,kc->bkhw", ycc, mi).clamp(0, 255) / 255
    return rgb[:, :, :h, :w]


def gaussian_noise(img: Tensor, sigma: Tensor, gray: Tensor, z: Tensor, z_gray: Tensor | None) -> Tensor:
    """random_add_gaussian_noise_pt(clip=True) with the draws given (degradations.py:569-605,665-676)."""
    b = img.size(0)
    s = sigma.view(b, 1, 1, 1)
    noise = z * s / 255.0
    if float(gray.sum()) > 0:
        g = gray.view(b, 1, 1, 1)
        noise = noise * (1 - g) + (z_gray.view(1, 1, *z_gray.shape[-2:]) * s / 255.0) * g
    return torch.clamp(img + noise, 0, 1)


def _vals(img_q: Tensor) -> Tensor:
    v = [2 ** math.ceil(math.log2(len(torch.unique(img_q[i])))) for i in range(img_q.size(0))]
    return img_q.new_tensor(v).view(-1, 1, 1, 1)


def rgb_to_gray(img: Tensor) -> Tensor:
    r, g, b = img.unbind(1)
    return (0.2989 * r + 0.587 * g + 0.114 * b).unsqueeze(1)  # torchvision rgb_to_grayscale


def poisson_noise(img: Tensor, scale: Tensor, gray: Tensor, counts_color: Tensor | None = None,
                  counts_gray: Tensor | None = None, generator=None):
    """random_add_poisson_noise_pt(clip=True) (degradations.py:738-786,851-862).  Returns
    (out, counts_color, counts_gray): the Poisson draws used (drawn here when not given)."""
    b = img.size(0)
    any_gray = float(gray.sum()) > 0
    g = gray.view(b, 1, 1, 1)
    noise_gray = None
    if any_gray:
        ig = torch.clamp((rgb_to_gray(img) * 255.0).round(), 0, 255) / 255.0
        vg = _vals(ig)
        if counts_gray is None:
            counts_gray = torch.poisson(ig * vg, generator=generator)
        noise_gray = (counts_gray / vg - ig).expand(b, 3, *img.shape[-2:])
    iq = torch.clamp((img * 255.0).round(), 0, 255) / 255.0
    vc = _vals(iq)
    if counts_color is None:
        counts_color = torch.poisson(iq * vc, generator=generator)
    noise = counts_color / vc - iq
    if any_gray:
        noise = noise * (1 - g) + noise_gray * g
    return torch.clamp(img + noise * scale.view(b, 1, 1, 1), 0, 1), counts_color, counts_gray


def resize(img: Tensor, mode: str, scale_factor=None, size=None) -> Tensor:
    return F.interpolate(img, scale_factor=scale_factor, size=size, mode=mode)


class Pool:
    """otf._dequeue_and_enqueue (otf.py:37-90), literally: full-pool gather by the given permutation."""

    def __init__(self, queue_size: int):
        self.queue_size, self.ptr, self.lr, self.gt = queue_size, 0, None, None

    def step(self, lq: Tensor, gt: Tensor, perm: Tensor | None):
        b = lq.size(0)
        if self.lr is None:
            assert self.queue_size % b == 0
            self.lr = torch.zeros(self.queue_size, *lq.shape[1:])
            self.gt = torch.zeros(self.queue_size, *gt.shape[1:])
        if self.ptr == self.queue_size:
            self.lr, self.gt = self.lr[perm], self.gt[perm]
            lq_d, gt_d = self.lr[:b].clone(), self.gt[:b].clone()
            self.lr[:b], self.gt[:b] = lq.clone(), gt.clone()
            return lq_d, gt_d
        self.lr[self.ptr:self.ptr + b], self.gt[self.ptr:self.ptr + b] = lq.clone(), gt.clone()
        self.ptr += b
        return lq, gt


def degrade(gt: Tensor, kernel1: Tensor, kernel2: Tensor, sinc_kernel: Tensor, plan: dict, scale: int,
            fields: dict | None = None, ds: dict | None = None, gen: torch.Generator | None = None):
    """otf.feed_data up to (and including) the random crop (otf.py:105-257).  `plan` holds the host-side
    decisions; the per-sample scalars (sigma*/pscale*/gray*/jpeg_q*) and the random `fields`
    ("z1","zg1","cc1","cg1","z2","zg2","cc2","cg2") are taken from plan/fields when present, otherwise drawn
    from `gen` with the ranges in `ds` in EXACTLY the order the reference draws them from the torch RNG
    (degradations.py:654-662,593-601,840-848,766-781; otf.py:150-152,229-238).  Returns
    (lq, gt_crop, completed plan, fields) so the CUDA path can replay the same numbers."""
    plan, fields = dict(plan), dict(fields or {})
    gen = gen or torch.Generator().manual_seed(int(plan.get("seed", 0)))
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)  # noqa: E731
    b, _, ori_h, ori_w = gt.shape

    def noise(out, i, sfx):
        gauss = plan[f"gauss{i}"]
        key = f"sigma{i}" if gauss else f"pscale{i}"
        if key not in plan:
            lo, hi = ds["noise_range" + sfx] if gauss else ds["poisson_scale_range" + sfx]
            plan[key] = (torch.rand(b, generator=gen) * (hi - lo) + lo).numpy()
            plan[f"gray{i}"] = (torch.rand(b, generator=gen) < ds["gray_noise_prob" + sfx]).float().numpy()
        any_gray = float(np.asarray(plan[f"gray{i}"]).sum()) > 0
        if gauss:
            if any_gray and f"zg{i}" not in fields:
                fields[f"zg{i}"] = torch.randn(out.shape[-2:], generator=gen)
            if f"z{i}" not in fields:
                fields[f"z{i}"] = torch.randn(out.shape, generator=gen)
            return gaussian_noise(out, t(plan[key]), t(plan[f"gray{i}"]), fields[f"z{i}"], fields.get(f"zg{i}"))
        out, cc, cg = poisson_noise(out, t(plan[key]), t(plan[f"gray{i}"]), fields.get(f"cc{i}"), fields.get(f"cg{i}"),
                                    generator=gen)
        fields[f"cc{i}"] = cc
        if cg is not None:
            fields[f"cg{i}"] = cg
        return out

    def jq(key, rkey):
        if key not in plan:
            lo, hi = ds[rkey]
            plan[key] = torch.zeros(b).uniform_(lo, hi, generator=gen).numpy()
        return t(plan[key])

    out = filter2d(gt, kernel1)
    out = resize(out, plan["mode1"], scale_factor=plan["scale1"])
    out = noise(out, 1, "")
    out = jpeg(torch.clamp(out, 0, 1), jq("jpeg_q1", "jpeg_range"))
    if plan["blur2"]:
        out = filter2d(out, kernel2)
    out = resize(out, plan["mode2"], size=(int(ori_h / scale * plan["scale2"]), int(ori_w / scale * plan["scale2"])))
    out = noise(out, 2, "2")
    final = (ori_h // scale, ori_w // scale)
    if plan["sinc_first"]:
        out = filter2d(resize(out, plan["mode3"], size=final), sinc_kernel)
        out = jpeg(torch.clamp(out, 0, 1), jq("jpeg_q2", "jpeg_range2"))
    else:
        out = jpeg(torch.clamp(out, 0, 1), jq("jpeg_q2", "jpeg_range2"))
        out = filter2d(resize(out, plan["mode3"], size=final), sinc_kernel)
    lq = torch.clamp((out * 255.0).round(), 0, 255) / 255.0
    ps, top, left = plan["patch_size"], plan["top"], plan["left"]
    lq = lq[:, :, top:top + ps, left:left + ps]
    gt = gt[:, :, top * scale:(top + ps) * scale, left * scale:(left + ps) * scale]
    return lq.contiguous(), gt.contiguous(), plan, fields


def apply_augment(gt: Tensor, lq: Tensor, scale: int, plan: dict):
    """apply_augment (neosr/data/augmentations.py:219-310) with every draw given by `plan`
    (neosr_b200.data.augmentations.draw_augment_plan): up-sample LQ (antialias), the mix operations in the
    reference's order, down-sample (bicubic antialias)."""
    gt, lq = gt.clone(), lq.clone()
    if scale > 1:
        lq = torch.clamp(F.interpolate(lq, scale_factor=scale, mode=plan["up_mode"], antialias=True), 0, 1)
    for op in plan["ops"]:
        perm = torch.as_tensor(np.asarray(op["perm"]), dtype=torch.long) if "perm" in op else None
        if op["op"] == "mixup":  # augmentations.py:14-33
            lam = op["lam"]
            g_ = gt[perm]
            gt, lq = lam * gt + (1 - lam) * g_, lam * lq + (1 - lam) * g_
        elif op["op"] == "cutmix":  # 36-61
            x1, y1, x2, y2 = op["box"]
            g_, l_ = gt[perm], lq[perm]
            gt[:, :, x1:x2, y1:y2] = g_[:, :, x1:x2, y1:y2]
            lq[:, :, x1:x2, y1:y2] = l_[:, :, x1:x2, y1:y2]
        elif op["op"] == "resizemix":  # 64-124
            x1, y1, x2, y2 = op["box"]
            g_, l_ = gt.clone()[perm], lq.clone()[perm]
            gt[:, :, y1:y2, x1:x2] = torch.clamp(F.interpolate(g_, (y2 - y1, x2 - x1), mode="bicubic", antialias=True), 0, 1)
            lq[:, :, y1:y2, x1:x2] = torch.clamp(F.interpolate(l_, (y2 - y1, x2 - x1), mode="bicubic", antialias=True), 0, 1)
        else:  # cutblur 127-166
            x1, y1, x2, y2 = op["box"]
            lq[:, :, x1:x2, y1:y2] = gt[:, :, x1:x2, y1:y2]
    if scale > 1:
        lq = torch.clamp(F.interpolate(lq, scale_factor=1 / scale, mode="bicubic", antialias=True), 0, 1)
    return gt, lq
