"""Drop-in hook: make the reference's own registries resolve to the B200 implementations.

neosr discovers plugins by scanning its package folders and registering by `__name__`
(neosr/archs/__init__.py:17-27, losses/__init__.py:14-22, models/__init__.py:13-22); a duplicate
name trips an assert (utils/registry.py:51-55).  `install_into_neosr()` therefore *replaces* the
`_obj_map` entries after the reference's scan has run, so an unmodified
`python train.py -opt options/train_swinir.toml` builds these modules for `network_g.type`,
`*_opt.type` and `model_type = "image"`.
"""
from __future__ import annotations

from .registry import ARCH_REGISTRY, LOSS_REGISTRY, MODEL_REGISTRY


def install_into_neosr(archs: bool = True, losses: bool = True, models: bool = True) -> dict:
    """Returns {registry: [names overridden]}.  Import neosr first (it must be on sys.path)."""
    from neosr.utils import registry as ref  # noqa: PLC0415
    from . import archs as _a, losses as _l, models as _m  # noqa: F401, PLC0415  (populate our registries)
    done: dict = {}
    for enabled, ours, theirs in ((archs, ARCH_REGISTRY, ref.ARCH_REGISTRY), (losses, LOSS_REGISTRY, ref.LOSS_REGISTRY),
                                  (models, MODEL_REGISTRY, ref.MODEL_REGISTRY)):
        if not enabled:
            continue
        names = []
        for name, obj in ours:
            theirs._obj_map[name] = obj  # override or add
            names.append(name)
        done[theirs._name] = names
    return done
