"""Import the *live* reference (muslll/neosr at /root/reference) on CPU.

TEST INFRASTRUCTURE, build-container only: /root/reference does not exist on the
GPU box.  Used by oracle/make_golden.py and tests/test_oracle_vs_reference.py to
pin the oracle against the reference's own modules and real step methods.

The shim does what SURVEY.md §8c lists: `-opt` argv stub (neosr parses options at
import time, archs/arch_util.py:12-27), stub `pywt`/`lmdb` modules, seeded
random-weight VGG19 (pretrained weights need network), cuda->cpu for the VGG
normalisation buffers (vgg_arch.py:168-174).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path

def _find_ref_root() -> Path:
    """$NEOSR_REFERENCE, else the read-only mount of the build container, else the staged copy that travels to the
    GPU box (baseline/_ref, written by baseline/install_ref.py; git-ignored)."""
    env = os.environ.get("NEOSR_REFERENCE")
    if env:
        return Path(env)
    mount = Path("/root/reference")
    if (mount / "neosr" / "archs" / "swinir_arch.py").exists():
        return mount
    return Path(__file__).resolve().parent.parent / "baseline" / "_ref"


REF_ROOT = _find_ref_root()

_TOML = """
name = "shim"
model_type = "image"
scale = {scale}
manual_seed = 1024
[datasets.train]
type = "paired"
patch_size = 64
batch_size = 4
[network_g]
type = "compact"
[train.optim_g]
type = "adan_sf"
lr = 1e-3
schedule_free = true
[logger]
total_iter = 100
"""

_state = {"scale": None}


def available() -> bool:
    return (REF_ROOT / "neosr" / "archs" / "swinir_arch.py").exists()


def activate(scale: int = 4):
    """Make `import neosr...` work. The scale is baked in at first import (module-level
    `upscale` in every arch); later calls must use the same scale or pass it as a kwarg."""
    if _state["scale"] is not None:
        return
    d = Path(tempfile.mkdtemp(prefix="neosr_shim_"))
    (d / "opt.toml").write_text(_TOML.format(scale=scale))
    sys.argv = [sys.argv[0] if sys.argv else "shim", "-opt", str(d / "opt.toml")]
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    for m in ("pywt", "lmdb"):
        try:
            __import__(m)
        except ImportError:
            sys.modules.setdefault(m, types.ModuleType(m))
    _state["scale"] = scale


def build_network(opt: dict):
    activate()
    from neosr.archs import build_network as bn  # noqa: PLC0415
    return bn(dict(opt))


def build_vgg_perceptual(seed_params: dict | None, loss_weight: float = 0.5, criterion: str = "chc"):
    """Reference vgg_perceptual_loss with OUR seeded VGG19 weights injected (None: torchvision's random init)."""
    activate()
    import torch  # noqa: PLC0415
    import torchvision.models.vgg as tvvgg  # noqa: PLC0415

    orig_vgg19 = tvvgg.vgg19
    orig_tensor = torch.tensor

    def fake_vgg19(weights=None, **kw):  # noqa: ARG001
        return orig_vgg19(weights=None)

    def cpu_tensor(*a, **kw):
        if str(kw.get("device", "")).startswith("cuda"):
            kw["device"] = "cpu"
        return orig_tensor(*a, **kw)

    from neosr.archs import vgg_arch  # noqa: PLC0415
    vgg_arch.vgg.vgg19 = fake_vgg19
    torch.tensor = cpu_tensor
    try:
        from neosr.losses.vgg_perceptual_loss import vgg_perceptual_loss  # noqa: PLC0415
        mod = vgg_perceptual_loss(loss_weight=loss_weight, criterion=criterion)
    finally:
        torch.tensor = orig_tensor
        vgg_arch.vgg.vgg19 = orig_vgg19
    sd = mod.vgg.state_dict()
    for k, v in (seed_params or {}).items():
        assert k in sd and tuple(sd[k].shape) == tuple(v.shape), k
        sd[k].copy_(v)
    return mod


def make_image_model(net_g, *, cri_pix=None, cri_perceptual=None, optim_kw=None, ema=0.999, scale=4,
                     net_d=None, cri_gan=None, optim_d_kw=None, device="cpu", eco=None, sam_init=None):
    """`object.__new__(image)` with the attributes `closure`/`optimize_parameters` read
    (image.py:73-230), so the reference's REAL step methods run on CPU."""
    activate()
    import torch  # noqa: PLC0415
    from torch.optim.swa_utils import AveragedModel, get_ema_multi_avg_fn  # noqa: PLC0415

    from neosr.models.image import image  # noqa: PLC0415
    from neosr.optimizers import adan_sf  # noqa: PLC0415

    m = object.__new__(image)
    m.opt = {"dist": False, "rank": 0, "world_size": 1, "scale": scale, "num_gpu": 1,
             "datasets": {"train": {}}, "train": {}, "path": {}}
    m.device = torch.device(device)
    m.is_train = True
    net_g = net_g.to(m.device)
    cri_pix = cri_pix.to(m.device) if cri_pix is not None else None
    cri_perceptual = cri_perceptual.to(m.device) if cri_perceptual is not None else None
    net_d = net_d.to(m.device) if net_d is not None else None
    cri_gan = cri_gan.to(m.device) if cri_gan is not None else None
    m.net_g, m.net_d = net_g, None
    m.optimizers, m.schedulers = [], []
    kw = dict(optim_kw or {})
    m.optimizer_g = adan_sf([p for p in net_g.parameters() if p.requires_grad], **kw)
    m.optimizers.append(m.optimizer_g)
    m.sf_optim_g, m.sf_optim_d = kw.get("schedule_free", True), None
    m.ema = ema
    if ema > 0:
        m.net_g_ema = AveragedModel(net_g, multi_avg_fn=get_ema_multi_avg_fn(ema), device=m.device)
    m.sam, m.sam_init = None, -1
    m.use_amp, m.amp_dtype = False, torch.float16
    m.gradscaler_g = torch.amp.GradScaler("cuda", enabled=False)
    m.eco, m.match_lq_colors, m.wavelet_guided, m.wavelet_init = False, False, False, 0
    m.n_accumulated, m.accum_iters, m.gradclip = 0, 1, True
    m.cri_pix, m.cri_perceptual = cri_pix, cri_perceptual
    for n in ("cri_mssim", "cri_consistency", "cri_dists", "cri_gan", "cri_ldl", "cri_ff", "cri_gw"):
        setattr(m, n, None)
    m.scale, m.aug, m.aug_prob, m.patch_size = scale, None, None, 64
    if eco:  # image.py:136-146
        m.eco, m.eco_schedule = True, eco.get("schedule", "sigmoid")
        m.eco_iters, m.eco_init, m.pretrain = eco.get("iters", 80000), eco.get("init", 15000), eco.get("pretrain")
    if sam_init is not None:  # image.py:90-91, 322-330
        from neosr.optimizers import fsam  # noqa: PLC0415
        m.sam, m.sam_init = "fsam", sam_init
        m.sam_optimizer_g = fsam([p for p in net_g.parameters() if p.requires_grad], adan_sf, rho=0.5, sigma=1, lmbda=0.9,
                                 adaptive=True, **kw)
        # the reference steps `sam_optimizer_g.base_optimizer` under SAM and `optimizer_g` before sam_init: two separate
        # adan_sf states over the same parameters (image.py:309-330); the oracle keeps ONE, so start SAM at iteration 0
        m.optimizer_g = m.sam_optimizer_g.base_optimizer
        m.optimizers = [m.optimizer_g]
    net_g.train()
    m.optimizer_g.train()
    if net_d is not None:  # image.py:40-46, 356-372
        dkw = dict(optim_d_kw or kw)
        m.net_d, m.cri_gan = net_d, cri_gan
        m.optimizer_d = adan_sf(list(net_d.parameters()), **dkw)
        m.optimizers.append(m.optimizer_d)
        m.sf_optim_d = dkw.get("schedule_free", True)
        m.gradscaler_d = torch.amp.GradScaler("cuda", enabled=False)
        net_d.train()
        m.optimizer_d.train()
    return m
