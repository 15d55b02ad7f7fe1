"""Validation-time host helpers (image.py:792-925): tensor -> uint8 image, file output, metrics.  Inside a neosr checkout
these resolve to neosr's own `tensor2img` / `imwrite` / `calculate_metric` (validation-time CPU code the drop-in does not
replace); standalone, minimal equivalents cover PSNR so validation still runs."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch


def tensor2img(t: torch.Tensor, rgb2bgr: bool = True) -> np.ndarray:
    """neosr/utils/img_util.py:60-129 for one [1,C,H,W] / [C,H,W] tensor in [0,1] -> HWC uint8 (BGR by default)."""
    t = t.squeeze(0).float().detach().cpu().clamp_(0, 1)
    a = t.numpy().transpose(1, 2, 0)
    if a.shape[2] == 1:
        a = a[:, :, 0]
    elif rgb2bgr:
        a = a[:, :, ::-1]
    return (a * 255.0).round().astype(np.uint8)


def imwrite(img: np.ndarray, path: str) -> None:
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    try:
        import cv2
        if not cv2.imwrite(path, img):
            raise OSError("Failed in writing images.")
    except ImportError:  # no OpenCV: raw dump next to the intended file
        np.save(path + ".npy", img)


def _to_y(img: np.ndarray) -> np.ndarray:  # BGR uint8 -> Y of YCbCr (BT.601), as neosr/metrics/metric_util.py:25-44
    f = img.astype(np.float32) / 255.0
    return ((f @ np.array([24.966, 128.553, 65.481], dtype=np.float32)) + 16.0)[..., None]


def calculate_psnr(img: np.ndarray, img2: np.ndarray, crop_border: int = 4, test_y_channel: bool = False, **kw) -> float:
    """neosr/metrics/calculate.py:16-66."""
    assert img.shape == img2.shape, f"Image shapes are different: {img.shape}, {img2.shape}."
    if crop_border:
        img, img2 = img[crop_border:-crop_border, crop_border:-crop_border, ...], img2[crop_border:-crop_border, crop_border:-crop_border, ...]
    if test_y_channel and img.ndim == 3 and img.shape[2] == 3:
        img, img2 = _to_y(img), _to_y(img2)
    mse = np.mean((img.astype(np.float64) - img2.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else float(10.0 * np.log10(255.0 * 255.0 / mse))


def calculate_metric(data: dict, opt: dict) -> float:
    try:
        from neosr.metrics import calculate_metric as ref_metric  # inside a neosr checkout: the reference's registry
        return ref_metric(data, opt)
    except ImportError:
        opt = dict(opt)
        kind = opt.pop("type")
        if kind in {"calculate_psnr", "psnr"}:
            return calculate_psnr(data["img"], data["img2"], **opt)
        raise NotImplementedError(f"metric {kind!r} needs neosr.metrics (not importable here); calculate_psnr is built in")
