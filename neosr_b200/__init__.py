"""neosr_b200 — B200-native (sm_100a) implementation of the neosr training-step hot path.

Drop-in surface (mirrors muslll/neosr): `ARCH_REGISTRY`, `LOSS_REGISTRY`, `MODEL_REGISTRY`,
`build_network`, `build_loss`, `build_model`, and `install_into_neosr()` which overrides the
reference's own registry entries so an unmodified `train.py -opt x.toml` runs on these kernels.
All compute goes through libneosr_b200.so (C ABI in include/neosr_b200.h); there is no
PyTorch/CPU fallback.
"""
from .registry import ARCH_REGISTRY, LOSS_REGISTRY, MODEL_REGISTRY  # noqa: F401

__version__ = "0.1.0"


def build_network(opt: dict):
    from .archs import build_network as _b
    return _b(opt)


def build_loss(opt: dict):
    from .losses import build_loss as _b
    return _b(opt)


def build_model(opt: dict):
    from .models import build_model as _b
    return _b(opt)


def install_into_neosr():
    from .plugin import install_into_neosr as _i
    return _i()
