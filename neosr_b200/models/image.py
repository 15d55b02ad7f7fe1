"""`image` model: the training step on the B200 kernels — drop-in for neosr/models/image.py.

train.py calls, per iteration, `feed_data(batch)` -> `optimize_parameters(it)` ->
`update_learning_rate(it, warmup_iter=)`, and on cadence `get_current_log()`,
`get_current_learning_rate()`, `save(epoch, it)` (train.py:264-332).  This class keeps those
methods, the option keys they read (image.py:73-230) and the log keys they produce
(`l_g_pix`, `l_g_percep`, `l_g_total`, ...), but runs ONE explicit pipeline instead of an
autograd graph:

    G fprop (saving activations) -> fused loss value+grad kernels -> G backward straight into
    one flat fp32 gradient buffer -> [NCCL all-reduce of that buffer, DDP-style] ->
    one fused kernel: clip_grad_norm_(1.0) + adan_sf/AdamW + EMA.

Differences from the reference that are deliberate:
  * multi-GPU works (the reference's DDP path cannot start); gradients are averaged across
    ranks exactly as DDP would, and the loss log is averaged over ranks only when read;
  * the loss log and the NaN check are materialised lazily (at `get_current_log()`), which
    removes the >= 2 host syncs per iteration (image.py:611, base.py:522-524).
"""
from __future__ import annotations

import copy
import os
import random
import time
from collections import OrderedDict
from pathlib import Path
from typing import Any

import numpy as np
import torch
from torch import Tensor
from torch.optim.swa_utils import AveragedModel, get_ema_multi_avg_fn

from .. import ops
from ..archs import build_network
from ..archs.arch_util import set_default_scale
from ..data.augmentations import draw_augment_plan, run_augment_plan
from ..dist import allreduce_mean_, broadcast_params_
from ..losses import build_loss
from ..optimizers import AdamW, adan_sf, fsam
from ..registry import MODEL_REGISTRY


def host_default_device(cls):
    """The reference's entry point sets `torch.set_default_device("cuda")` for the whole process (train.py:165); this
    package is written against torch's normal default (host tensors unless a device is named: CPU generators, pinned
    staging buffers, option scalars).  Every method the training loop can reach therefore runs with the default device
    pinned back to the CPU; device tensors are always created with an explicit `device=`."""
    import functools

    def wrap(fn):
        @functools.wraps(fn)
        def inner(*a, **kw):
            with torch.device("cpu"):
                return fn(*a, **kw)
        return inner

    for name, attr in list(vars(cls).items()):
        if callable(attr) and not isinstance(attr, (staticmethod, classmethod, type)) and (name == "__init__" or not name.startswith("__")):
            setattr(cls, name, wrap(attr))
    return cls


@MODEL_REGISTRY.register()
@host_default_device
class image:
    def __init__(self, opt: dict[str, Any]) -> None:
        self.opt = opt
        if not torch.cuda.is_available():
            raise RuntimeError("neosr_b200.image: a CUDA device (B200, sm_100a) is required; there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.is_train = opt.get("is_train", True)
        self.optimizers: list = []
        self.schedulers: list = []
        self._graph_mode = False
        self.is_train_mode = False
        self.ema = -1
        self.log_dict: dict = {}
        set_default_scale(opt.get("scale", 4), self.is_train)

        self.net_g = build_network(opt["network_g"]).to(self.device)
        self.net_d = None
        if opt.get("network_d") is not None:  # image.py:40-46
            self.net_d = build_network(opt["network_d"]).to(self.device)
            if not hasattr(self.net_d, "engine_backward"):
                raise NotImplementedError("neosr_b200.image: network_d must be an engine-backed arch (unet)")
        path = opt.get("path", {}) or {}
        if path.get("pretrain_network_g"):
            self.load_network(self.net_g, path["pretrain_network_g"], path.get("param_key_g"),
                              path.get("strict_load_g", True))
        if self.net_d is not None and path.get("pretrain_network_d"):
            self.load_network(self.net_d, path["pretrain_network_d"], path.get("param_key_d", "params"),
                              path.get("strict_load_d", True))
        self.dist = bool(opt.get("dist", False))
        self.world_size = int(opt.get("world_size", 1))
        if self.dist and self.world_size > 1:
            broadcast_params_(self.net_g.parameters())  # identical replicas, as DDP does when wrapping
            if self.net_d is not None:
                broadcast_params_([*self.net_d.parameters(), *self.net_d.buffers()])
        if self.is_train:
            self.init_training_settings()

    # ------------------------------------------------------------------ setup (image.py:73-372)
    def init_training_settings(self) -> None:
        train_opt = self.opt["train"]
        self.ema = train_opt.get("ema", -1)
        if self.ema > 0:
            self.net_g_ema = AveragedModel(self.net_g, multi_avg_fn=get_ema_multi_avg_fn(self.ema), device=self.device)
        if train_opt.get("wavelet_guided"):
            raise NotImplementedError("neosr_b200.image: train.wavelet_guided is outside the built hot path (SURVEY.md §8f.4; "
                                      "the reference's wavelet transform needs PyWavelets)")
        self.match_lq_colors = bool(train_opt.get("match_lq_colors", False))  # image.py:130, 451-463, 484-485
        # opt-in step variants (image.py:90-91, 136-146): F-SAM double closure and ECO centroid targets
        self.sam = train_opt.get("sam", None)
        self.sam_init = train_opt.get("sam_init", -1)
        if self.sam is not None and self.sam not in {"FSAM", "fsam"}:
            raise NotImplementedError(f"SAM type {self.sam} not supported yet.")
        if self.sam is not None and self.opt.get("network_d") is not None:
            raise NotImplementedError("neosr_b200.image: F-SAM together with a discriminator is not built (the reference's second "
                                      "closure accumulates the discriminator gradients of both passes)")
        self.eco = bool(train_opt.get("eco", False))
        self.eco_schedule = train_opt.get("eco_schedule", "sigmoid")
        self.eco_iters = train_opt.get("eco_iters", 80000)
        self.eco_init = train_opt.get("eco_init", 15000)
        self.pretrain = (self.opt.get("path") or {}).get("pretrain_network_g")
        # image.py:117-127: `use_amp` (+ `bfloat16`).  Here mixed precision = every tcgen05 contraction issues ONE bf16 pass
        # (fp32 accumulate, fp32 storage; NSR_ENGINE_BF16) instead of the three of the fp32-parity split: no autocast
        # copies and no GradScaler - bf16 has fp32's range, and gradients are formed and stored in fp32 - so the
        # reference's scaler state is not mirrored.  float16 autocast is not built.
        self.use_amp = self.opt.get("use_amp", False) is True
        if self.use_amp:
            if self.opt.get("bfloat16", False) is not True:
                raise NotImplementedError("neosr_b200.image: use_amp needs bfloat16 = true (float16 autocast is not built)")
            ops.DEFAULT_ENGINE = "bf16"  # process-wide, like torch.autocast in the reference's closure
        elif ops.DEFAULT_ENGINE == "bf16":
            ops.DEFAULT_ENGINE = "auto"  # a model without use_amp built after one with it gets the fp32-parity engine back
        ds = self.opt.get("datasets", {}).get("train", {})
        self.accum_iters = ds.get("accumulate", 1) or 1
        if self.accum_iters != 1:
            raise NotImplementedError("neosr_b200.image: accumulate != 1 (ill-defined in the reference, SURVEY.md §3.2)")
        self.aug = ds.get("augmentation")  # image.py:114-115
        self.aug_prob = ds.get("aug_prob")
        ps = ds.get("patch_size")
        if self.aug is not None and ps is not None and ps % 4 != 0:  # image.py:275-277
            raise ValueError("The patch_size value must be a multiple of 4 while using augmentations.")
        seed = int(self.opt.get("manual_seed", 1024) or 1024) + int(self.opt.get("rank", 0))
        self._aug_rng, self._aug_pyrandom = np.random.default_rng([seed, 2]), random.Random(seed + 2)
        self._aug_rng_dev = np.random.default_rng([seed, 3])
        self.n_accumulated = 0
        self._ema_updates = 0  # host mirror of net_g_ema.n_averaged (avoids a device read per step)
        self._ema_params = None
        # CUDA-graph replay of the step (opt-out: `cuda_graph = false` in the option file); needs a
        # deterministic launch sequence, i.e. no DropPath draws, and the device-scalar optimizer path
        # (read from the built network, not from the option text: hat_* defaults to drop_path_rate 0.1 too)
        dp = max((float(getattr(m, "drop_prob", 0.0) or 0.0) for m in self.net_g.modules()), default=0.0)
        sf_types = {"adan_sf", "Adan_SF", "AdamW", "adamw"}  # fused optimizers with device-resident per-step scalars
        self._graph_mode = (bool(self.opt.get("cuda_graph", True)) and dp == 0.0 and self.sam is None and not self.eco
                            and train_opt["optim_g"].get("type") in sf_types
                            and (self.net_d is None or (train_opt.get("optim_d") or {}).get("type") in sf_types))
        self._graphs, self._graph_logs, self._eager_steps = None, None, 0
        self._lq_static = self._gt_static = None
        self.scale = self.opt.get("scale", 4)
        self.patch_size = ds.get("patch_size")
        self.gradclip = train_opt.get("grad_clip", True)

        def mk(key):
            o = train_opt.get(key)
            if not o:
                return None
            o = dict(o)
            if key == "perceptual_opt" and "allow_random_init" not in o and self.opt.get("vgg_random_init"):
                o["allow_random_init"] = True
            return build_loss(o).to(self.device)

        self.cri_pix = mk("pixel_opt")
        self.cri_perceptual = mk("perceptual_opt")
        self.cri_gan = mk("gan_opt")
        self.cri_mssim = mk("mssim_opt")
        self.cri_consistency = mk("consistency_opt")
        for k in ("dists_opt", "ldl_opt", "ff_opt", "gw_opt"):
            if train_opt.get(k):
                raise NotImplementedError(f"neosr_b200.image: train.{k} not built yet")
        if self.cri_pix is None and self.cri_mssim is None and self.cri_perceptual is None:
            raise ValueError("Both pixel/mssim and perceptual losses are None. Please enable at least one.")
        optim_d = train_opt.get("optim_d")  # image.py:259-275
        if self.net_d is None and optim_d is not None:
            raise ValueError("Please set a discriminator in network_d or disable optim_d.")
        if self.net_d is not None and optim_d is None:
            raise ValueError("Please set an optimizer for the discriminator or disable network_d.")
        if self.net_d is not None and self.cri_gan is None:
            raise ValueError("Discriminator needs GAN to be enabled.")
        if self.net_d is None and self.cri_gan is not None:
            raise ValueError("GAN requires a discriminator to be set.")
        self.setup_optimizers()
        self.setup_schedulers()
        self.net_g.train()
        if self.sf_optim_g:
            self.optimizer_g.train()
        if self.net_d is not None:
            self.net_d.train()
            if self.sf_optim_d:
                self.optimizer_d.train()

    @staticmethod
    def _make_optimizer(params, o: dict):
        o = dict(o)
        optim_type = o.pop("type")
        if optim_type in {"Adan_SF", "adan_sf"}:
            if "schedule_free" not in o:
                raise ValueError("The option 'schedule_free' must be in the config file.")
            return adan_sf(params, **o)
        if optim_type in {"AdamW", "adamw"}:
            return AdamW(params, **o)
        raise NotImplementedError(f"neosr_b200.image: optimizer {optim_type} not built (adan_sf, AdamW are)")

    def setup_optimizers(self) -> None:  # image.py:340-372
        o = self.opt["train"]["optim_g"]
        self.sf_optim_g = o.get("schedule_free", False)
        g_params = [p for p in self.net_g.parameters() if p.requires_grad]
        self.optimizer_g = self._make_optimizer(g_params, o)
        self.optimizers.append(self.optimizer_g)
        if getattr(self, "sam", None) is not None:  # image.py:309-330: F-SAM around a second instance of the base optimizer
            kw = dict(o)
            t = kw.pop("type")
            base = adan_sf if t in {"Adan_SF", "adan_sf"} else (AdamW if t in {"AdamW", "adamw"} else None)
            if base is None:
                raise NotImplementedError(f"SAM not supported by optimizer {t} yet.")
            self.sam_optimizer_g = fsam(g_params, base, rho=0.5, sigma=1, lmbda=0.9, adaptive=True, **kw)
        if self.net_d is not None:
            o = self.opt["train"]["optim_d"]
            self.sf_optim_d = o.get("schedule_free", False)
            self.optimizer_d = self._make_optimizer(list(self.net_d.parameters()), o)
            self.optimizers.append(self.optimizer_d)

    def setup_schedulers(self) -> None:  # base.py:174-198
        """`[train.scheduler]`: MultiStepLR / CosineAnnealing over every optimizer, as the reference builds them (they
        also give each param group its `initial_lr`, which the linear warm-up reads).  Unknown types are an error."""
        so = self.opt["train"].get("scheduler")
        if so is None:
            return
        so = dict(so)
        kind = so.pop("type")
        if kind in {"MultiStepLR", "multisteplr"}:
            cls = torch.optim.lr_scheduler.MultiStepLR
        elif kind in {"CosineAnnealing", "cosineannealing"}:
            cls = torch.optim.lr_scheduler.CosineAnnealingLR
        else:
            raise NotImplementedError(f"Scheduler {kind} is not implemented yet.")
        for o in self.optimizers:
            self.schedulers.append(cls(o, **so))

    # ------------------------------------------------------------------ the hot path
    @torch.no_grad()
    def feed_data(self, data: dict) -> None:  # image.py:374-391
        lq, gt = data["lq"], data.get("gt")
        if self.is_train and self.aug is not None and gt is not None and not (len(self.aug) == 1 and "none" in self.aug):
            # image.py:380-391: mixup / cutmix / resizemix / cutblur on the device
            lq = lq.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
            gt = gt.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
            plan = draw_augment_plan(gt.size(0), gt.size(2), gt.size(3), self.scale, self.aug, self.aug_prob, self._aug_rng,
                                     self._aug_pyrandom, self._aug_rng_dev)
            gt, lq = run_augment_plan(gt, lq, self.scale, plan)
        if self._graph_mode:
            # CUDA-graph replay needs fixed input addresses: copy the batch into static device buffers
            if self._lq_static is None or self._lq_static.shape != lq.shape or (gt is not None and self._gt_static.shape != gt.shape):
                self._lq_static = torch.empty(lq.shape, dtype=torch.float32, device=self.device)
                self._gt_static = torch.empty(gt.shape, dtype=torch.float32, device=self.device) if gt is not None else None
                self._graphs = None  # shapes changed: re-capture
                self._eager_steps = 0
            self._lq_static.copy_(lq, non_blocking=True)
            if gt is not None:
                self._gt_static.copy_(gt, non_blocking=True)
            self.lq, self.gt = self._lq_static, self._gt_static
            return
        self.lq = lq.to(self.device, non_blocking=True)
        if gt is not None:
            self.gt = gt.to(self.device, non_blocking=True)

    def _eco_input(self, current_iter: int) -> Tensor:
        """eco_strategy (image.py:393-425): the network input is the LQ centroid (1-a) down(G(lq)) + a lq and the target
        the GT centroid (1-a) G(lq) + a gt (self.gt is overwritten for the rest of the step, as in the reference)."""
        import math
        if self.eco_schedule == "sigmoid":
            a = 1 / (1 + math.exp(-1 * (10 * (current_iter / self.eco_iters - 0.25))))
        else:
            a = min(current_iter / self.eco_iters, 1.0)
        net_output, _ = self.net_g.engine_forward(self.lq, save=False)
        self.gt = ops.axpby(net_output, 1.0 - a, self.gt, a)
        lq_scaled = ops.resize_aa(net_output, "bicubic", scale_factor=1 / self.scale)  # clamps to [0, 1]
        return ops.axpby(lq_scaled, 1.0 - a, self.lq, a)

    def _forward_backward(self, current_iter: int = 0) -> OrderedDict:
        """G fprop -> fused loss value+grad kernels -> G backward into the flat gradient buffer.
        Returns the loss scalars as device tensors (image.py:448-531)."""
        net = self.net_g
        ps = net.param_set()
        ps.ensure_grads(self.device)
        lq = self.lq
        if self.eco and current_iter <= self.eco_iters and not (current_iter < self.eco_init and self.pretrain is None):
            lq = self._eco_input(current_iter)
        out, saved = net.engine_forward(lq, save=True)
        self.output = out
        total = torch.zeros(1, dtype=torch.float32, device=self.device)
        logs = OrderedDict()
        dout = None
        if self.cri_pix is not None:
            v, g = self.cri_pix.value_and_grad(out, self.gt, True, total)
            logs["l_g_pix"] = v
            dout = g
        if self.cri_mssim is not None:  # image.py:479-482
            v, g = self.cri_mssim.value_and_grad(out, self.gt, True, total)
            logs["l_g_mssim"] = v
            dout = g if dout is None else ops.axpby(dout, 1.0, g, 1.0, out=dout)
        if self.cri_consistency is not None:  # image.py:484-489
            tgt = self.gt
            if self.match_lq_colors:  # colours / luma are matched to the up-sampled LQ instead of the GT (451-463)
                tgt = ops.clamp(ops.resize_aa(self.lq, "bicubic", scale_factor=self.scale), 1.0 / 255.0, 1.0)
            v, g = self.cri_consistency.value_and_grad(out, tgt, True, total)
            logs["l_g_consistency"] = v
            dout = g if dout is None else ops.axpby(dout, 1.0, g, 1.0, out=dout)
        if self.cri_perceptual is not None:
            v, g = self.cri_perceptual.value_and_grad(out, self.gt, True, total)
            logs["l_g_percep"] = v
            dout = g if dout is None else ops.axpby(dout, 1.0, g, 1.0, out=dout)
        if self.cri_gan is not None:
            # generator's adversarial term (image.py:516-520): D is frozen, only d(loss)/d(output) flows back
            pred, sd = self.net_d.engine_forward(out, save=True)
            v, g = self.cri_gan.value_and_grad(pred, True, False, True, total)
            logs["l_g_gan"] = v
            g = self.net_d.engine_backward(sd, g, param_grads=False, need_dx=True)
            del sd
            dout = g if dout is None else ops.axpby(dout, 1.0, g, 1.0, out=dout)
        logs["l_g_total"] = total
        net.engine_backward(saved, dout)
        del saved
        if self.net_d is not None:
            self._discriminator_backward(out, logs)
        return logs

    def _discriminator_backward(self, out: Tensor, logs: OrderedDict) -> None:
        """Real and fake passes of net_d (image.py:547-596); the fake pass accumulates onto the real pass's
        gradients.  Each backward runs before the next forward (spectral-norm state is per forward)."""
        d = self.net_d
        pred, sd = d.engine_forward(self.gt, save=True)
        l_real, g = self.cri_gan.value_and_grad(pred, True, True, True, None)
        logs["l_d_real"], logs["out_d_real"] = l_real, pred.mean()
        d.engine_backward(sd, g, param_grads=True, accumulate=False, need_dx=False)
        del sd
        pred, sd = d.engine_forward(out, save=True)  # `out` carries no graph: the reference's .detach()
        l_fake, g = self.cri_gan.value_and_grad(pred, False, True, True, None)
        logs["l_d_fake"], logs["out_d_fake"] = l_fake, pred.mean()
        d.engine_backward(sd, g, param_grads=True, accumulate=True, need_dx=False)
        logs["l_d_total"] = (l_real + l_fake) / 2

    def _ema_arg(self):
        if self.ema <= 0:
            return None
        if self._ema_params is None:
            self._ema_params = [e.detach() for e, p in zip(self.net_g_ema.module.parameters(), self.net_g.parameters())
                                if p.requires_grad]
        return (self._ema_params, self.ema, self._ema_updates == 0)

    @torch.no_grad()
    def optimize_parameters(self, current_iter: int) -> None:  # image.py:427-662
        ps = self.net_g.param_set()
        clip = 1.0 if self.gradclip else None
        multi = self.dist and self.world_size > 1
        if self._graph_mode and self._graphs is None and self._eager_steps >= 2:
            self._capture_graphs(clip)
        if self._graph_mode and self._graphs is not None:
            # graph 1: G forward, losses, G backward [, D real / fake forward + backward] | gradient all-reduce(s) |
            # graph 2: grad-norm + fused optimizer + EMA of G [, of D]
            g_fb, g_opt, n_fb, n_opt = self._graphs
            g_fb.replay()
            if multi:
                allreduce_mean_(ps.flat_grad)
                if self.net_d is not None:
                    allreduce_mean_(self.net_d.param_set().flat_grad)
            self.optimizer_g.prepare(clip_max_norm=clip, ema=self._ema_arg(), to_device=True)
            if self.net_d is not None:
                self.optimizer_d.prepare(clip_max_norm=clip, ema=None, to_device=True)
            g_opt.replay()
            self.optimizer_g.bump_versions()
            if self.net_d is not None:
                self.optimizer_d.bump_versions()
            ops._count(n_fb + n_opt)
            logs = self._graph_logs
        elif self.sam is not None and current_iter >= self.sam_init:
            # fsam.step (fsam.py:82-95): gradients at w -> climb to w + e(w) -> gradients there -> back to w, base optimizer
            # step.  No generator clip under SAM (image.py:533-537).
            self._forward_backward(current_iter)
            if multi:
                allreduce_mean_(ps.flat_grad)
            ps.attach_grads()
            self.sam_optimizer_g.first_step()
            logs = self._forward_backward(current_iter)
            if multi:
                allreduce_mean_(ps.flat_grad)
            ps.attach_grads()
            self.sam_optimizer_g.second_step(clip_max_norm=None, ema=self._ema_arg())
            self._eager_steps += 1
        else:
            logs = self._forward_backward(current_iter)
            if multi:
                allreduce_mean_(ps.flat_grad)  # the one collective of the step (DDP-style gradient averaging)
            ps.attach_grads()
            self.optimizer_g.step(clip_max_norm=clip, ema=self._ema_arg())
            if self.net_d is not None:  # image.py:598-608, 645
                psd = self.net_d.param_set()
                if multi:
                    allreduce_mean_(psd.flat_grad)
                psd.attach_grads()
                self.optimizer_d.step(clip_max_norm=clip, ema=None)
            self._eager_steps += 1
        if self.ema > 0:
            self.net_g_ema.n_averaged += 1
            self._ema_updates += 1
        self._pending_logs = logs

    def _capture_graphs(self, clip) -> None:
        """Capture the step into two CUDA graphs (forward+loss+backward | grad-norm+optimizer+EMA) so the
        ~2000 launches of a step cost no host time; the gradient all-reduce (N > 1) and the 72-byte
        upload of the optimizer's per-step scalars stay between/around the replays."""
        ps = self.net_g.param_set()
        ps.invalidate_packed()  # the weight re-pack kernels must be part of the captured step
        if self.net_d is not None:
            self.net_d.param_set().invalidate_packed()
        torch.cuda.synchronize()
        l0 = ops.LAUNCHES
        g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_fb):
            self._graph_logs = self._forward_backward()
        n_fb = ops.LAUNCHES - l0
        ps.attach_grads()
        # plan (tables, device-resident scalars) without advancing the schedule twice: prepare() is
        # re-run before every replay, the capture only needs the launch sequence and stable pointers
        opts = [(self.optimizer_g, self._ema_arg())]
        if self.net_d is not None:
            self.net_d.param_set().attach_grads()
            opts.append((self.optimizer_d, None))
        for o, ema in opts:
            if isinstance(o, AdamW):  # torch-style per-parameter `step` tensors: rewind the one this planning pass adds
                o.prepare(clip_max_norm=clip, ema=ema, to_device=True)
                for st in o.state.values():
                    st["step"] -= 1
                continue
            saved = [dict(step=g.get("step"), weight_sum=g["weight_sum"], lr_max=g["lr_max"]) for g in o.param_groups]
            o.prepare(clip_max_norm=clip, ema=ema, to_device=True)
            for g, sv in zip(o.param_groups, saved):
                g["weight_sum"], g["lr_max"] = sv["weight_sum"], sv["lr_max"]
                if sv["step"] is None:
                    g.pop("step", None)
                else:
                    g["step"] = sv["step"]
        torch.cuda.synchronize()
        l0 = ops.LAUNCHES
        g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_opt):
            for o, _ in opts:
                o.launch()
        n_opt = ops.LAUNCHES - l0
        self._graphs = (g_fb, g_opt, n_fb, n_opt)

    def update_learning_rate(self, current_iter: int, warmup_iter: int = -1) -> None:  # base.py:229-254
        if current_iter > 0 and self.n_accumulated == 0:
            for s in self.schedulers:
                s.step()
        if current_iter < warmup_iter:
            # linear warm-up from the lr the scheduler recorded (base.py:203-252: `_get_init_lr` reads
            # param_group["initial_lr"], which exists only when [train.scheduler] is set - same KeyError here)
            for opt in self.optimizers:
                for g in opt.param_groups:
                    g["lr"] = g["initial_lr"] / warmup_iter * current_iter

    def get_current_learning_rate(self):
        return [g["lr"] for g in self.optimizers[0].param_groups]

    def get_current_log(self) -> dict:
        """Materialise the (device-resident) loss scalars; average over ranks if distributed
        (what base.reduce_loss_dict intends, base.py:498-526)."""
        logs = getattr(self, "_pending_logs", None)
        if logs:
            keys = list(logs)
            vec = torch.cat([logs[k].view(1) for k in keys])
            if self.dist and self.world_size > 1:
                allreduce_mean_(vec)
            vals = vec.tolist()
            if any(v != v for v in vals):
                raise ValueError("NaN found, aborting training. Make sure you're using a proper learning rate.")
            self.log_dict = OrderedDict(zip(keys, vals))
        return self.log_dict

    # ------------------------------------------------------------------ checkpoints (base.py:281-496)
    def get_bare_model(self, net):
        return net

    def save_network(self, net, net_label: str, current_iter, param_key: str = "params") -> None:
        current_iter = "latest" if current_iter == -1 else current_iter
        models_dir = Path(self.opt["path"]["models"])
        models_dir.mkdir(parents=True, exist_ok=True)
        sd = OrderedDict()
        for k, v in net.state_dict().items():
            if k == "n_averaged":
                continue
            k = k[7:] if k.startswith("module.") else k
            sd[k] = v.detach().cpu()
        if self.sf_optim_g and self.is_train:
            self.optimizer_g.eval()
        torch.save({param_key: sd}, models_dir / f"{net_label}_{current_iter}.pth")
        if self.sf_optim_g and self.is_train:
            self.optimizer_g.train()

    def save_training_state(self, epoch: int, current_iter: int) -> None:
        if current_iter == -1:
            return
        d = Path(self.opt["path"]["training_states"])
        d.mkdir(parents=True, exist_ok=True)
        state = {"epoch": epoch, "iter": current_iter, "optimizers": [o.state_dict() for o in self.optimizers],
                 "schedulers": [s.state_dict() for s in self.schedulers]}
        torch.save(state, d / f"{current_iter}.state")

    def save(self, epoch: int, current_iter: int) -> None:  # image.py:932-942
        if self.opt.get("rank", 0) != 0:
            return
        if self.ema > 0:
            self.save_network(self.net_g_ema, "net_g", current_iter)
        else:
            self.save_network(self.net_g, "net_g", current_iter)
        if self.net_d is not None:  # image.py:939-940
            self.save_network(self.net_d, "net_d", current_iter)
        self.save_training_state(epoch, current_iter)

    def load_network(self, net, load_path, param_key: str | None = None, strict: bool = True) -> None:
        load_net = torch.load(load_path, map_location="cpu", weights_only=True)
        if param_key is None:
            for k in ("params-ema", "params_ema", "params"):
                if k in load_net:
                    param_key = k
                    break
        if param_key in load_net:
            load_net = load_net[param_key]
        load_net = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in load_net.items())
        net.load_state_dict(load_net, strict=strict)

    def resume_training(self, resume_state: dict) -> None:  # base.py:477-496
        for i, o in enumerate(resume_state["optimizers"]):
            self.optimizers[i].load_state_dict(o)
        for i, s in enumerate(resume_state["schedulers"]):
            self.schedulers[i].load_state_dict(s)

    # ------------------------------------------------------------------ validation (image.py:664-925)
    def _infer(self, x: Tensor) -> Tensor:
        net = self.net_g_ema if (self.is_train_mode and self.ema > 0) else self.net_g
        return net(x.to(self.device, dtype=torch.float32).contiguous())

    @torch.no_grad()
    def test(self) -> None:
        """EMA-weights inference on self.lq: whole image (`val.tile = -1`) or the reference's partition scheme with a
        16-pixel overlap (image.py:664-784).  The forward is the same engine kernels in save=False mode."""
        self.tile = (self.opt.get("val") or {}).get("tile", -1)
        scale = self.opt["scale"]
        self.is_train_mode = getattr(self, "optimizer_g", None) is not None
        sf = self.is_train_mode and getattr(self, "sf_optim_g", False)
        if sf:
            self.optimizer_g.eval()
        if self.is_train_mode and self.ema > 0:
            self.net_g_ema.eval()  # also drops the packed-weight images: the fused optimizer updates EMA weights in place
        else:
            self.net_g.eval()
        try:
            if self.tile == -1:
                self.output = self._infer(self.lq)
            else:
                _, C, h, w = self.lq.size()
                nh, nw = h // self.tile + 1, w // self.tile + 1
                pad_h = (nh - h % nh) % nh
                pad_w = (nw - w % nw) % nw
                img = self.lq
                img = torch.cat([img, torch.flip(img, [2])], 2)[:, :, : h + pad_h, :]
                img = torch.cat([img, torch.flip(img, [3])], 3)[:, :, :, : w + pad_w]
                _, _, H, W = img.size()
                sh, sw, shave = H // nh, W // nw, 16
                ral, row = H // sh, W // sw
                out = torch.zeros(1, C, H * scale, W * scale, device=self.device)
                for i in range(ral):
                    for j in range(row):
                        top = slice(i * sh - (shave if i > 0 else 0), (i + 1) * sh + (shave if i < ral - 1 else 0))
                        left = slice(j * sw - (shave if j > 0 else 0), (j + 1) * sw + (shave if j < row - 1 else 0))
                        o = self._infer(img[..., top, left].contiguous())
                        _top = slice(0, sh * scale) if i == 0 else slice(shave * scale, (shave + sh) * scale)
                        _left = slice(0, sw * scale) if j == 0 else slice(shave * scale, (shave + sw) * scale)
                        out[..., i * sh * scale:(i + 1) * sh * scale, j * sw * scale:(j + 1) * sw * scale] = o[..., _top, _left]
                self.output = out[:, :, 0:H * scale - pad_h * scale, 0:W * scale - pad_w * scale]
        finally:
            self.net_g.train()
            if sf:
                self.optimizer_g.train()

    def get_current_visuals(self) -> OrderedDict:  # image.py:925-931
        out = OrderedDict(lq=self.lq.detach().cpu(), result=self.output.detach().cpu())
        if getattr(self, "gt", None) is not None:
            out["gt"] = self.gt.detach().cpu()
        return out

    def validation(self, dataloader, current_iter, tb_logger, save_img: bool = True) -> None:  # base.py:64-77
        if self.opt.get("dist"):
            if self.opt.get("rank", 0) == 0:  # image.py:786-790
                self.nondist_validation(dataloader, current_iter, tb_logger, save_img)
        else:
            self.nondist_validation(dataloader, current_iter, tb_logger, save_img)

    def nondist_validation(self, dataloader, current_iter, tb_logger, save_img: bool = True) -> None:  # image.py:792-901
        from ._validation import calculate_metric, imwrite, tensor2img
        was_train, self.is_train = self.is_train, False  # no augmentation / graph buffers during validation
        graph_mode, self._graph_mode = self._graph_mode, False
        vopt = self.opt.get("val") or {}
        ds_opt = getattr(dataloader.dataset, "opt", {}) or {}
        dataset_name, dataset_type = ds_opt.get("name", "val"), ds_opt.get("type", "paired")
        with_metrics = dataset_type != "single" and vopt.get("metrics") is not None
        if with_metrics:
            self.metric_results = dict.fromkeys(vopt["metrics"].keys(), 0.0)
            if not hasattr(self, "best_metric_results"):
                self.best_metric_results = {}
            self.best_metric_results.setdefault(dataset_name, {
                m: {"better": c.get("better", "higher"), "val": float("-inf") if c.get("better", "higher") == "higher" else float("inf"),
                    "iter": -1} for m, c in vopt["metrics"].items()})
        n = 0
        try:
            for val_data in dataloader:
                lq_path = val_data.get("lq_path", [f"img{n}"])
                name = Path(lq_path[0] if isinstance(lq_path, (list, tuple)) else lq_path).stem
                self.feed_data(val_data)
                self.test()
                vis = self.get_current_visuals()
                sr = tensor2img(vis["result"])
                data = {"img": sr}
                if "gt" in vis:
                    data["img2"] = tensor2img(vis["gt"])
                    self.gt = None
                self.lq = self.output = None
                if vopt.get("save_img", save_img) and (self.opt.get("path") or {}).get("visualization"):
                    root = Path(self.opt["path"]["visualization"])
                    if self.opt.get("is_train", True):
                        f = root / name / f"{name}_{current_iter}.png"
                    else:
                        f = root / dataset_name / f"{name}_{vopt.get('suffix') or self.opt.get('name', 'b200')}.png"
                    imwrite(sr, str(f))
                if with_metrics:
                    for m, c in vopt["metrics"].items():
                        self.metric_results[m] += calculate_metric(data, c)
                n += 1
        finally:
            self.is_train, self._graph_mode = was_train, graph_mode
        if with_metrics and n:
            for m in self.metric_results:
                self.metric_results[m] /= n
                best = self.best_metric_results[dataset_name][m]
                v = self.metric_results[m]
                if (best["better"] == "higher" and v >= best["val"]) or (best["better"] != "higher" and v <= best["val"]):
                    best["val"], best["iter"] = v, current_iter
                if tb_logger:
                    tb_logger.add_scalar(f"metrics/{dataset_name}/{m}", v, current_iter)
