"""Host-side blur-kernel synthesis for the OTF pipeline.

The reference builds three 21x21 kernels per sample in DataLoader workers
(neosr/data/otf_dataset.py:189-246, neosr/data/degradations.py:107-512): float64 numpy, a few hundred
flops each — they are the *inputs* of the device pipeline (`nsr_filter2d`) and stay on the host here
too.  Same draw order from the same two generators (python `random` for the discrete choices, a numpy
Generator for the uniforms), so a seeded run reproduces the reference's kernels.
"""
from __future__ import annotations

import math
import random as _random

import numpy as np
from scipy import special


def _grid(k: int) -> np.ndarray:  # degradations.py:44-62 (mesh_grid)
    ax = np.arange(-k // 2 + 1.0, k // 2 + 1.0)
    xx, yy = np.meshgrid(ax, ax)
    return np.stack([xx, yy], -1)


def _sigma_matrix(sx: float, sy: float, theta: float, isotropic: bool) -> np.ndarray:  # degradations.py:24-41
    if isotropic:
        return np.array([[sx**2, 0.0], [0.0, sx**2]])
    u = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
    return u @ np.diag([sx**2, sy**2]) @ u.T


def _quad(k: int, sx: float, sy: float, theta: float, isotropic: bool) -> np.ndarray:
    g = _grid(k)
    inv = np.linalg.inv(_sigma_matrix(sx, sy, theta, isotropic))
    return np.sum((g @ inv) * g, 2)


def bivariate_gaussian(k, sx, sy, theta, isotropic=True):  # degradations.py:107-134
    ker = np.exp(-0.5 * _quad(k, sx, sy, theta, isotropic))
    return ker / ker.sum()


def bivariate_generalized_gaussian(k, sx, sy, theta, beta, isotropic=True):  # degradations.py:137-170
    ker = np.exp(-0.5 * np.power(_quad(k, sx, sy, theta, isotropic), beta))
    return ker / ker.sum()


def bivariate_plateau(k, sx, sy, theta, beta, isotropic=True):  # degradations.py:173-210
    ker = np.reciprocal(np.power(_quad(k, sx, sy, theta, isotropic), beta) + 1)
    return ker / ker.sum()


def circular_lowpass_kernel(cutoff: float, k: int, pad_to: int = 0) -> np.ndarray:  # degradations.py:477-512
    assert k % 2 == 1, "Kernel size must be an odd number."
    c = (k - 1) / 2
    yy, xx = np.mgrid[0:k, 0:k]
    r = np.sqrt((yy - c) ** 2 + (xx - c) ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        ker = cutoff * special.j1(cutoff * r) / (2 * np.pi * r)
    ker[(k - 1) // 2, (k - 1) // 2] = cutoff**2 / (4 * np.pi)
    ker = ker / ker.sum()
    if pad_to > k:
        p = (pad_to - k) // 2
        ker = np.pad(ker, ((p, p), (p, p)))
    return ker


def random_mixed_kernels(kernel_list, kernel_prob, k, sigma_x_range, sigma_y_range, rotation_range, betag_range,
                         betap_range, rng: np.random.Generator, pyrandom=_random) -> np.ndarray:
    """degradations.py:379-471 with `noise_range=None` (what otf_dataset.py passes)."""
    kind = pyrandom.choices(kernel_list, kernel_prob)[0]
    iso = kind in {"iso", "generalized_iso", "plateau_iso"}
    assert k % 2 == 1, "Kernel size must be an odd number."
    sx = rng.uniform(sigma_x_range[0], sigma_x_range[1])
    if iso:
        sy, rot = sx, 0.0
    else:
        sy = rng.uniform(sigma_y_range[0], sigma_y_range[1])
        rot = rng.uniform(rotation_range[0], rotation_range[1])
    if kind in {"iso", "aniso"}:
        ker = bivariate_gaussian(k, sx, sy, rot, iso)
    else:
        br = betag_range if kind.startswith("generalized") else betap_range
        beta = rng.uniform(br[0], 1) if rng.uniform() < 0.5 else rng.uniform(1, br[1])
        fn = bivariate_generalized_gaussian if kind.startswith("generalized") else bivariate_plateau
        ker = fn(k, sx, sy, rot, beta, iso)
    return ker / ker.sum()


KERNEL_RANGE = [2 * v + 1 for v in range(3, 11)]  # otf_dataset.py:110 (7..21, hard-coded)


def _one(opt: dict, sfx: str, rng, pyrandom) -> np.ndarray:
    k = pyrandom.choice(KERNEL_RANGE)
    if rng.uniform() < opt.get("sinc_prob" + sfx):
        omega = rng.uniform(np.pi / 3, np.pi) if k < 13 else rng.uniform(np.pi / 5, np.pi)
        ker = circular_lowpass_kernel(omega, k, pad_to=False)
    else:
        ker = random_mixed_kernels(opt.get("kernel_list" + sfx), opt.get("kernel_prob" + sfx), k,
                                   opt.get("blur_sigma" + sfx), opt.get("blur_sigma" + sfx), [-math.pi, math.pi],
                                   opt.get("betag_range" + sfx), opt.get("betap_range" + sfx), rng, pyrandom)
    p = (21 - k) // 2
    return np.pad(ker, ((p, p), (p, p)))


def synth_kernels(opt: dict, rng: np.random.Generator, pyrandom=_random):
    """The three kernels of one sample, fp32 [21,21] each (otf_dataset.py:189-246)."""
    k1 = _one(opt, "", rng, pyrandom)
    k2 = _one(opt, "2", rng, pyrandom)
    if rng.uniform() < opt.get("final_sinc_prob"):
        k = pyrandom.choice(KERNEL_RANGE)
        sinc = circular_lowpass_kernel(rng.uniform(np.pi / 3, np.pi), k, pad_to=21)
    else:
        sinc = np.zeros((21, 21))
        sinc[10, 10] = 1.0  # pulse: no blur
    return k1.astype(np.float32), k2.astype(np.float32), sinc.astype(np.float32)


# the [degradations] table shared by options/train_hat_otf.toml and options/train_realplksr_otf.toml
TEMPLATE_DEGRADATIONS = dict(  # options/train_hat_otf.toml / train_realplksr_otf.toml [degradations]
    resize_prob=[0.3, 0.4, 0.3], resize_range=[0.5, 1.5], gaussian_noise_prob=0.2, noise_range=[0, 2],
    poisson_scale_range=[0.05, 0.25], gray_noise_prob=0.1, jpeg_range=[40, 95], second_blur_prob=0.4,
    resize_prob2=[0.3, 0.4, 0.3], resize_range2=[0.3, 1.5], gaussian_noise_prob2=0.2, noise_range2=[0, 2],
    poisson_scale_range2=[0.05, 0.1], gray_noise_prob2=0.1, jpeg_range2=[35, 95],
    blur_kernel_size=7, kernel_list=["iso", "aniso", "generalized_iso", "generalized_aniso", "plateau_iso", "plateau_aniso"],
    kernel_prob=[0.45, 0.25, 0.12, 0.03, 0.12, 0.03], sinc_prob=0.1, blur_sigma=[0.2, 3], betag_range=[0.5, 4],
    betap_range=[1, 2], blur_kernel_size2=9,
    kernel_list2=["iso", "aniso", "generalized_iso", "generalized_aniso", "plateau_iso", "plateau_aniso"],
    kernel_prob2=[0.45, 0.25, 0.12, 0.03, 0.12, 0.03], sinc_prob2=0.1, blur_sigma2=[0.2, 1.5], betag_range2=[0.5, 4],
    betap_range2=[1, 2], final_sinc_prob=0.8)
