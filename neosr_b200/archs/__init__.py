"""Arch factory with the reference's contract (neosr/archs/__init__.py:14-34):
`build_network({"type": name, **kwargs}) -> nn.Module`."""
from __future__ import annotations

from copy import deepcopy

from ..registry import ARCH_REGISTRY
from . import swinir_arch  # noqa: F401  (registers swinir_*)
from . import compact_arch  # noqa: F401
from . import esrgan_arch  # noqa: F401
from . import unet_arch  # noqa: F401
from . import realplksr_arch  # noqa: F401
from . import hat_arch  # noqa: F401
from . import vgg_arch  # noqa: F401


def build_network(opt: dict):
    opt = deepcopy(opt)
    network_type = opt.pop("type")
    return ARCH_REGISTRY.get(network_type)(**opt)
