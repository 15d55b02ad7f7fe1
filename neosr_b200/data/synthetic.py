"""Synthetic paired LQ/GT crops: GT = seeded uniform noise quantised to 8 bit (real
data is uint8/255, neosr/utils/img_util.py:176-180); LQ = antialiased bicubic downsample, clamped
and quantised.  Used by bench.py and the tests; there is no network access for real datasets."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.utils.data import Dataset


def synth_pair(index: int, lq_size: int = 64, scale: int = 4, seed: int = 1024):
    g = torch.Generator(device="cpu").manual_seed(seed * 1_000_003 + index)
    gt = torch.rand(1, 3, lq_size * scale, lq_size * scale, generator=g)
    gt = torch.round(gt * 255) / 255
    lq = F.interpolate(gt, scale_factor=1 / scale, mode="bicubic", antialias=True).clamp(0, 1)
    lq = torch.round(lq * 255) / 255
    return lq[0], gt[0]


class SyntheticPairedDataset(Dataset):
    """Returns the reference's paired-dataset item dict (paired_dataset.py:168)."""

    def __init__(self, length: int = 1024, lq_size: int = 64, scale: int = 4, seed: int = 1024):
        self.length, self.lq_size, self.scale, self.seed = length, lq_size, scale, seed

    def __len__(self) -> int:
        return self.length

    def __getitem__(self, i: int) -> dict:
        lq, gt = synth_pair(i, self.lq_size, self.scale, self.seed)
        return {"lq": lq, "gt": gt, "lq_path": f"synthetic/{i}", "gt_path": f"synthetic/{i}"}


def structured_gt(seed: int, b: int, h: int, w: int) -> torch.Tensor:
    """Structured synthetic GT: low-pass noise + step edges + a flat patch, on 8-bit levels,
    so the JPEG quantiser, the Poisson level count and the clamps are all exercised."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, 3, h, w, generator=g)
    x = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(x, (4, 4, 4, 4), mode="reflect"), 9, 1)
    x = (x - x.amin((1, 2, 3), keepdim=True)) / (x.amax((1, 2, 3), keepdim=True) - x.amin((1, 2, 3), keepdim=True))
    x[:, :, h // 3:, w // 2:] = 1.0 - x[:, :, h // 3:, w // 2:]
    x[:, :, : h // 4, : w // 4] = 0.25
    x = x + 0.02 * torch.rand(b, 3, h, w, generator=g)
    return torch.round(x.clamp(0, 1) * 255) / 255
