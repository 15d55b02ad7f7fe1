"""Run the reference's REAL `otf.feed_data` on CPU and record every random decision and field it draws
(build container only; TEST INFRASTRUCTURE, see oracle/__init__.py).

The recording is turned into the `plan` / `fields` dictionaries that `oracle.otf.degrade` and
`neosr_b200.models.otf.run_plan` take, so both can be replayed on exactly what the reference did
(`object.__new__(otf)` with hand-set attributes; CPU gotchas 4 and 7 handled here).
"""
from __future__ import annotations

import random

import numpy as np
import torch

from oracle import ref_shim

from neosr_b200.data.degradations import TEMPLATE_DEGRADATIONS as DEGRADATIONS  # noqa: E402
from neosr_b200.data.synthetic import structured_gt  # noqa: E402,F401


def make_ref_otf(ds: dict, scale: int, queue_size: int):
    """`object.__new__(otf)` ready for the real feed_data on CPU."""
    ref_shim.activate(4)
    import neosr.utils.diffjpeg as dj
    dj.device = torch.device("cpu")
    import neosr.models.otf as om
    m = object.__new__(om.otf)
    m.opt = {"scale": scale, "datasets": {"train": dict(ds)}}
    m.is_train, m.device, m.scale = True, torch.device("cpu"), scale
    m.jpeger = dj.DiffJPEG(differentiable=False)
    m.queue_size, m.patch_size, m.aug, m.aug_prob = queue_size, ds["patch_size"], None, None
    return m, om


class Recorder:
    """Patches the RNG entry points otf.feed_data reaches and logs what they return."""

    def __init__(self, om, seed: int):
        self.om, self.seed = om, seed
        self.log: dict[str, list] = {k: [] for k in ("choices", "choice", "uniform", "randint", "rand", "randn",
                                                     "poisson", "uniform_", "randperm")}

    def __enter__(self):
        import neosr.data.transforms as tr
        om, log = self.om, self.log
        self._saved = (om.random, om.rng, om.filter2D, tr.random, torch.rand, torch.randn, torch.poisson,
                       torch.Tensor.uniform_, torch.randperm)
        pyr = random.Random(self.seed)
        nprng = np.random.default_rng(self.seed)
        gen = torch.Generator().manual_seed(self.seed)
        o_rand, o_randn, o_poisson, o_uniform_, o_randperm = torch.rand, torch.randn, torch.poisson, torch.Tensor.uniform_, torch.randperm

        class PyR:
            @staticmethod
            def choices(pop, w=None):
                r = pyr.choices(pop, w); log["choices"].append(r[0]); return r
            @staticmethod
            def choice(seq):
                r = pyr.choice(seq); log["choice"].append(r); return r
            @staticmethod
            def randint(a, b):
                r = pyr.randint(a, b); log["randint"].append(r); return r

        class NpR:
            @staticmethod
            def uniform(*a, **k):
                r = nprng.uniform(*a, **k); log["uniform"].append(float(r)); return r

        def rand(*a, **k):
            k.pop("device", None); r = o_rand(*a, generator=gen, **k); log["rand"].append(r.clone()); return r
        def randn(*a, **k):
            k.pop("device", None); r = o_randn(*a, generator=gen, **k); log["randn"].append(r.clone()); return r
        def poisson(lam, generator=None):
            r = o_poisson(lam, generator=gen); log["poisson"].append(r.clone()); return r
        def uniform_(self_, lo=0.0, hi=1.0, **k):
            r = o_uniform_(self_, lo, hi, generator=gen); log["uniform_"].append(r.clone()); return r
        def randperm(n, **k):
            k.pop("device", None); r = o_randperm(n, generator=gen); log["randperm"].append(r.clone()); return r

        f2d = om.filter2D
        om.random, om.rng, tr.random = PyR, NpR, PyR
        om.filter2D = lambda img, k: f2d(img.contiguous(), k)  # CPU reflection_pad2d keeps channels-last strides
        torch.rand, torch.randn, torch.poisson, torch.Tensor.uniform_, torch.randperm = rand, randn, poisson, uniform_, randperm
        return self

    def __exit__(self, *a):
        import neosr.data.transforms as tr
        om = self.om
        (om.random, om.rng, om.filter2D, tr.random, torch.rand, torch.randn, torch.poisson, torch.Tensor.uniform_,
         torch.randperm) = self._saved

    def plan_and_fields(self, ds: dict, patch_size: int):
        """Interpret the log in the order otf.feed_data draws (otf.py:105-257)."""
        L = {k: list(v) for k, v in self.log.items()}
        plan: dict = {"patch_size": patch_size, "seed": self.seed}
        fields: dict = {}

        def updown():
            t = L["choices"].pop(0)
            return 1 if t == "keep" else L["uniform"].pop(0)

        def noise(i, sfx):
            gauss = L["uniform"].pop(0) < ds["gaussian_noise_prob" + sfx]
            plan[f"gauss{i}"] = bool(gauss)
            lo, hi = ds["noise_range" + sfx] if gauss else ds["poisson_scale_range" + sfx]
            plan[f"sigma{i}" if gauss else f"pscale{i}"] = (L["rand"].pop(0) * (hi - lo) + lo).numpy()
            gray = (L["rand"].pop(0) < ds["gray_noise_prob" + sfx]).float()
            plan[f"gray{i}"] = gray.numpy()
            any_gray = float(gray.sum()) > 0
            if gauss:
                if any_gray:
                    fields[f"zg{i}"] = L["randn"].pop(0)
                fields[f"z{i}"] = L["randn"].pop(0)
            else:
                if any_gray:
                    fields[f"cg{i}"] = L["poisson"].pop(0)
                fields[f"cc{i}"] = L["poisson"].pop(0)

        plan["scale1"] = updown()
        plan["mode1"] = L["choice"].pop(0)
        noise(1, "")
        plan["jpeg_q1"] = L["uniform_"].pop(0).numpy()  # logged BEFORE quality_to_factor mutates it in place
        plan["blur2"] = bool(L["uniform"].pop(0) < ds["second_blur_prob"])
        plan["scale2"] = updown()
        plan["mode2"] = L["choice"].pop(0)
        noise(2, "2")
        plan["sinc_first"] = bool(L["uniform"].pop(0) < 0.5)
        plan["mode3"] = L["choice"].pop(0)
        plan["jpeg_q2"] = L["uniform_"].pop(0).numpy()
        plan["top"], plan["left"] = L["randint"].pop(0), L["randint"].pop(0)
        perm = L["randperm"].pop(0) if L["randperm"] else None
        assert not any(L[k] for k in ("choices", "choice", "uniform", "randint", "rand", "randn", "poisson", "uniform_")), L
        return plan, fields, perm


def run_reference(gt, k1, k2, sk, ds: dict, scale: int, seed: int, model=None, queue_size: int = 0):
    """One real otf.feed_data call.  Returns (lq, gt_crop, plan, fields, perm, model)."""
    ds = dict(ds)
    if model is None:
        model, _ = make_ref_otf(ds, scale, queue_size or (180 // gt.size(0)) * gt.size(0))
    import neosr.models.otf as om
    ps = ds["patch_size"]
    if not hasattr(model, "queue_lr"):  # pre-create the pool on CPU: the reference calls .cuda() (otf.py:57,65)
        model.queue_lr = torch.zeros(model.queue_size, 3, ps, ps)
        model.queue_gt = torch.zeros(model.queue_size, 3, ps * scale, ps * scale)
        model.queue_ptr = 0
    with Recorder(om, seed) as rec:
        model.feed_data({"gt": gt.clone(), "kernel1": k1.clone(), "kernel2": k2.clone(), "sinc_kernel": sk.clone()})
    plan, fields, perm = rec.plan_and_fields(ds, ps)
    return model.lq.clone(), model.gt.clone(), plan, fields, perm, model
