"""Model factory with the reference's contract (neosr/models/__init__.py:24-37)."""
from __future__ import annotations

from ..registry import MODEL_REGISTRY
from . import image  # noqa: F401  (registers `image`)
from . import otf  # noqa: F401  (registers `otf`)


def build_model(opt: dict):
    return MODEL_REGISTRY.get(opt["model_type"])(opt)
