"""Drop-in hook: make the reference's own registries resolve to the B200 implementations.

neosr discovers plugins by scanning its package folders and registering by `__name__`
(neosr/archs/__init__.py:17-27 lazily inside build_network, losses/__init__.py:14-22 and models/__init__.py:13-22 at
package import); a duplicate name trips an assert (utils/registry.py:51-55).  `install_into_neosr()` therefore first
makes every one of those scans happen (so the originals are registered and will not be registered again), then
*replaces* the `_obj_map` entries of the same name, so an unmodified `python train.py -opt x.toml` builds these
modules for `network_g.type` / `network_d.type`, `*_opt.type` and `model_type = "image" | "otf"`.

The maintainer-side change is one two-line file in a scanned folder:

    # neosr/models/zz_b200.py
    import neosr_b200
    neosr_b200.install_into_neosr()

`os.scandir` order is arbitrary, so the hook may run before image.py / otf.py were imported: it imports them itself.
"""
from __future__ import annotations

import importlib
import os
from pathlib import Path

from .registry import ARCH_REGISTRY, LOSS_REGISTRY, MODEL_REGISTRY


def _force_reference_scans() -> None:
    import neosr.archs as ref_archs  # noqa: PLC0415
    import neosr.losses  # noqa: F401, PLC0415  (scans *_loss.py at import)
    for f in sorted(Path(ref_archs.__file__).resolve().parent.glob("*_arch.py")):  # what build_network would import
        importlib.import_module(f"neosr.archs.{f.stem}")
    for name in ("image", "otf"):
        importlib.import_module(f"neosr.models.{name}")


def install_into_neosr(archs: bool = True, losses: bool = True, models: bool = True) -> dict:
    """Returns {registry: [names overridden]}.  `neosr` must be importable (on sys.path)."""
    from neosr.utils import registry as ref  # noqa: PLC0415
    from . import archs as _a, losses as _l, models as _m  # noqa: F401, PLC0415  (populate our registries)
    _force_reference_scans()
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
    if os.environ.get("NSR_PLUGIN_VERBOSE"):
        print("neosr_b200.install_into_neosr:", {k: len(v) for k, v in done.items()}, flush=True)
    return done
