"""Oracle: one `feed_data` + `optimize_parameters` iteration (TEST INFRASTRUCTURE).

Restates image.closure (neosr/models/image.py:427-625) and
image.optimize_parameters (627-662) for the configurations C1 (L1), C3 (L1 + VGG perceptual) and C2 (adding the U-Net
discriminator + BCE GAN loss, image.py:516-520, 547-608), accumulate = 1, AMP off, no SAM/ECO.
Autograd supplies the backward pass, exactly as in the reference.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
from torch import Tensor

from . import losses as L
from .optim import AdamWState, AdanSFState, EMAState, adamw_step, adan_sf_step, clip_grad_norm
from .swinir import SwinIRConfig, swinir_forward


class OracleTrainer:
    """Holds what `image.__init__`/`init_training_settings` hold: net_g params, the loss
    configuration, optimizer_g (adan_sf), EMA.  `net_fn(params, lq)` is the generator."""

    def __init__(self, params: dict, net_fn, *, pixel_weight: float | None = 1.0,
                 percep_weight: float | None = None, vgg_params: dict | None = None,
                 layer_weights: dict | None = None, optim: dict | None = None,
                 ema: float = 0.999, grad_clip: bool = True, disc: tuple | None = None,
                 gan_weight: float = 0.1, optim_d: dict | None = None, optim_type: str = "adan_sf",
                 mssim_weight: float | None = None, consistency_weight: float | None = None,
                 eco: dict | None = None, sam: dict | None = None, scale: int = 4, match_lq_colors: bool = False):
        self.names = list(params)
        self.params = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        self.net_fn = net_fn
        self.pixel_weight, self.percep_weight = pixel_weight, percep_weight
        self.mssim_weight, self.consistency_weight = mssim_weight, consistency_weight
        self.vgg_params, self.layer_weights = vgg_params, layer_weights
        self.grad_clip = grad_clip
        self.optim_type = optim_type  # adan_sf (templates' default) or AdamW (C5, base.py:154-155)
        state_cls = AdamWState if optim_type == "adamw" else AdanSFState
        self.opt = state_cls([self.params[k] for k in self.names], **(optim or {}))
        self.ema = EMAState([self.params[k] for k in self.names], ema) if ema and ema > 0 else None
        self.log_dict: OrderedDict = OrderedDict()
        self.last_grads: dict = {}
        # opt-in step variants: ECO (image.py:393-425; keys iters, init, schedule, pretrain) and F-SAM
        # (optimizers/fsam.py; keys init = sam_init; rho 0.5, sigma 1, lmbda 0.9, adaptive as image.py:322-330 builds it)
        self.eco, self.sam, self.scale = eco, sam, scale
        self.match_lq_colors = match_lq_colors  # image.py:451-463, 484-485
        self.sam_momentum: dict = {}
        # discriminator: (params, buffers) of oracle.unet; buffers are updated in place by each forward
        self.disc = None
        if disc is not None:
            dp, db = disc
            self.d_names = list(dp)
            self.d_params = {k: v.detach().clone().requires_grad_(True) for k, v in dp.items()}
            self.d_buffers = {k: v.detach().clone() for k, v in db.items()}
            self.disc = True
            self.gan_weight = gan_weight
            self.opt_d = AdanSFState([self.d_params[k] for k in self.d_names], **(optim_d or optim or {}))
            self.last_grads_d: dict = {}

    def feed_data(self, data: dict) -> None:  # image.py:374-391 (no augmentation)
        self.lq, self.gt = data["lq"], data["gt"]

    def _generator_output(self, current_iter: int):
        """image.closure 445-455 + eco_strategy 393-425: plain forward, or the ECO centroid pair (overwrites self.gt)."""
        e = self.eco
        if not e or current_iter > e.get("iters", 80000) or (current_iter < e.get("init", 15000) and not e.get("pretrain")):
            return self.net_fn(self.params, self.lq)
        import math

        import torch.nn.functional as F
        with torch.no_grad():
            if e.get("schedule", "sigmoid") == "sigmoid":
                a = 1 / (1 + math.exp(-1 * (10 * (current_iter / e.get("iters", 80000) - 0.25))))
            else:
                a = min(current_iter / e.get("iters", 80000), 1.0)
            net_output = self.net_fn(self.params, self.lq)
            self.gt = ((1 - a) * net_output) + (a * self.gt)
            lq_scaled = torch.clamp(F.interpolate(net_output, scale_factor=1 / self.scale, mode="bicubic", antialias=True), 0, 1)
            inp = ((1 - a) * lq_scaled) + (a * self.lq)
        return self.net_fn(self.params, inp)

    def optimize_parameters(self, current_iter: int = 0) -> None:
        s = self.sam
        if s and current_iter >= s.get("init", -1):
            # fsam.step (fsam.py:82-95): closure at w -> first_step (climb to w + e(w)) -> closure -> second_step (back to w,
            # base optimizer steps with the gradients taken at w + e(w)); the EMA follows the final weights
            if self.disc:
                raise NotImplementedError("oracle: F-SAM with a discriminator")
            grads = self._closure(current_iter)
            plist = [self.params[k] for k in self.names]
            sigma, lmbda, rho = s.get("sigma", 1.0), s.get("lmbda", 0.9), s.get("rho", 0.5)
            for k, p, g in zip(self.names, plist, grads):  # fsam.py:37-49
                grad = g.clone()
                if k not in self.sam_momentum:
                    self.sam_momentum[k] = grad
                else:
                    g -= self.sam_momentum[k] * sigma
                    self.sam_momentum[k] = self.sam_momentum[k] * lmbda + grad * (1 - lmbda)
            adaptive = s.get("adaptive", True)
            norm = torch.norm(torch.stack([((p.detach().abs() if adaptive else 1.0) * g).norm(p=2)
                                           for p, g in zip(plist, grads)]), p=2)  # fsam.py:97-110
            sc = rho / (norm + 1e-12)
            old = [p.detach().clone() for p in plist]
            with torch.no_grad():
                for p, g in zip(plist, grads):  # fsam.py:51-65
                    p.add_((torch.pow(p, 2) if adaptive else 1.0) * g * sc)
            grads = self._closure(current_iter)
            with torch.no_grad():
                for p, o in zip(plist, old):
                    p.copy_(o)
            (adamw_step if self.optim_type == "adamw" else adan_sf_step)(self.opt, grads)
            if self.ema is not None:
                self.ema.update(plist)
            return
        grads = self._closure(current_iter)
        plist = [self.params[k] for k in self.names]
        (adamw_step if self.optim_type == "adamw" else adan_sf_step)(self.opt, grads)  # image.py:642
        if self.disc:
            adan_sf_step(self.opt_d, self._dgrads)  # image.py:645
        if self.ema is not None:  # image.py:661-662
            self.ema.update(plist)

    def _closure(self, current_iter: int = 0) -> list:
        """image.closure (427-625): forward, loss stack, backward, clip; discriminator passes.  Returns the clipped
        generator gradients."""
        out = self._generator_output(current_iter)
        self.output = out
        total = torch.zeros(1)
        log = OrderedDict()
        if self.pixel_weight is not None:  # image.py:473-476
            l_pix = L.l1_loss(out, self.gt, self.pixel_weight)
            total = total + l_pix
            log["l_g_pix"] = l_pix
        if self.mssim_weight is not None:  # image.py:478-482
            l_ms = L.msssim_loss(out, self.gt, self.mssim_weight)
            total = total + l_ms
            log["l_g_mssim"] = l_ms
        if self.consistency_weight is not None:  # image.py:484-491
            tgt = self.gt
            if self.match_lq_colors:
                import torch.nn.functional as F
                tgt = torch.clamp(F.interpolate(self.lq, scale_factor=self.scale, mode="bicubic", antialias=True), 1 / 255, 1)
            l_co = L.consistency_loss(out, tgt, self.consistency_weight)
            total = total + l_co
            log["l_g_consistency"] = l_co
        if self.percep_weight is not None:  # image.py:491-494
            l_per = L.vgg_perceptual_loss(self.vgg_params, out, self.gt, self.percep_weight, self.layer_weights)
            total = total + l_per
            log["l_g_percep"] = l_per
        if self.disc:  # image.py:516-520 — net_d frozen (requires_grad False): gradient reaches net_g only
            from .unet import unet_forward
            frozen = {k: v.detach() for k, v in self.d_params.items()}
            l_gan = L.gan_loss(unet_forward(frozen, self.d_buffers, out, True), True, False, loss_weight=self.gan_weight)
            total = total + l_gan
            log["l_g_gan"] = l_gan
        log["l_g_total"] = total
        plist = [self.params[k] for k in self.names]
        grads = torch.autograd.grad(total, plist, allow_unused=True)
        grads = [torch.zeros_like(p) if g is None else g.contiguous().clone() for g, p in zip(grads, plist)]
        sam_on = bool(self.sam) and current_iter >= self.sam.get("init", -1)  # no generator clip under SAM (image.py:533-537)
        self.grad_norm = clip_grad_norm(grads, 1.0) if (self.grad_clip and not sam_on) else None  # image.py:533-544
        self.last_grads = dict(zip(self.names, [g.clone() for g in grads]))
        if self.disc:  # image.py:547-608
            real = unet_forward(self.d_params, self.d_buffers, self.gt, True)
            l_real = L.gan_loss(real, True, True)
            fake = unet_forward(self.d_params, self.d_buffers, out.detach(), True)
            l_fake = L.gan_loss(fake, False, True)
            log["l_d_real"], log["out_d_real"] = l_real, real.detach().mean()
            log["l_d_fake"], log["out_d_fake"] = l_fake, fake.detach().mean()
            log["l_d_total"] = (l_real + l_fake) / 2
            dlist = [self.d_params[k] for k in self.d_names]
            dgrads = [g.contiguous().clone() for g in torch.autograd.grad(l_real + l_fake, dlist)]
            if self.grad_clip:
                clip_grad_norm(dgrads, 1.0)
            self.last_grads_d = dict(zip(self.d_names, [g.clone() for g in dgrads]))
            self._dgrads = dgrads
        if torch.isnan(total).any():  # image.py:611-619
            raise ValueError("NaN found, aborting training.")
        self.log_dict = OrderedDict((k, float(v.detach().mean())) for k, v in log.items())
        return grads

    def get_current_log(self):
        return self.log_dict


def make_swinir_trainer(params: dict, cfg: SwinIRConfig, **kw) -> OracleTrainer:
    return OracleTrainer(params, lambda p, x: swinir_forward(p, cfg, x), **kw)


def displacement_report(p0: dict, ours: dict, ref32: dict, ref64: dict, noise_tol: float = 1e-2) -> dict:
    """Compare parameter displacements after k training steps (TEST INFRASTRUCTURE).

    adan's update is lr * m / sqrt(v): an element whose gradient is at round-off level (e.g. the key bias of an attention
    layer, whose true gradient is zero because softmax ignores a constant shift) still moves by ~lr per step, in a
    direction decided by noise - in the reference as much as anywhere.  Such elements are identified with the reference
    itself: where its fp32 and fp64 runs (`ref32`, `ref64`: name -> tensor) disagree by more than `noise_tol` of the step,
    fp32 round-off decides the step and the element says nothing about an implementation.  On the rest, the displacement
    of `ours` is compared with the fp32 reference's.  The key third of every `*.attn.qkv.bias` is excluded outright: its
    gradient is identically zero, so the fp32-vs-fp64 test only catches those of its elements where the oracle's own noise
    happened to disagree - the others would be scored on a coin toss (seen: 8 % on one qkv.bias under F-SAM after a change
    that only re-ordered fp32 sums).

    Returns {"worst": (rel-L2 error on the mask, name), "coverage": min over tensors of the fraction of elements kept,
    "coverage_all": fraction of ALL elements kept, "cos": cosine between the two displacement vectors over ALL elements
    (masked or not), "per_tensor": {name: (err, coverage)}}."""
    per, worst, cov_min = {}, (0.0, ""), 1.0
    dot = na = nb = 0.0
    kept = total = 0
    for k, init in p0.items():
        if k not in ours or not init.is_floating_point():
            continue
        i64 = init.double()
        d32, d64 = ref32[k].detach().double() - i64, ref64[k].detach().double() - i64
        do = ours[k].detach().double().cpu() - i64
        mask = (d32 - d64).abs() <= noise_tol * d64.abs()
        if k.endswith("attn.qkv.bias") and mask.dim() == 1 and mask.numel() % 3 == 0:
            c = mask.numel() // 3
            mask[c:2 * c] = False  # key bias: d softmax / d(constant shift of the scores) = 0
        cov = float(mask.double().mean())
        kept += int(mask.sum())
        total += mask.numel()
        err = float(((do - d32) * mask).norm() / (d32 * mask).norm().clamp_min(1e-300))
        per[k] = (err, cov)
        if err > worst[0]:
            worst = (err, k)
        cov_min = min(cov_min, cov)
        dot += float((do * d32).sum())
        na += float((do * do).sum())
        nb += float((d32 * d32).sum())
    return {"worst": worst, "coverage": cov_min, "coverage_all": kept / max(total, 1), "cos": dot / max((na * nb) ** 0.5, 1e-300), "per_tensor": per}
