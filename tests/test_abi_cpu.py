"""The C-ABI shared library builds for sm_100a, loads on a CPU-only box and exports every symbol
that include/neosr_b200.h declares (no compute calls without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    txt = (ROOT / "include" / "neosr_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nsr_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from neosr_b200 import _lib
    if not _lib.LIB_PATH.exists():
        _lib.build()
    dll = _lib.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(dll, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) <= set(names)
    assert dll.nsr_version() >= 100
    assert dll.nsr_last_error() is not None


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the ABI structs: pointer-sized fields, no implicit padding surprises."""
    from neosr_b200._lib import NsrAdanSF, NsrConv, NsrParamEntry, NsrWgrad
    assert ctypes.sizeof(NsrParamEntry) == 8 * 8 + 16
    # 15 ints/floats, pad to 8, 11 pointers, 5 ints (res_ld, aux_ld, pre_mode, sti_win, aux_mode) padded to 8, ws ptr + size
    assert ctypes.sizeof(NsrConv) == 15 * 4 + 4 + 11 * 8 + 24 + 16
    assert ctypes.sizeof(NsrWgrad) == 11 * 4 + 4 + 5 * 8 + 8 + 2 * 8
    assert ctypes.sizeof(NsrAdanSF) == 18 * 4


def test_pure_size_queries_need_no_gpu():
    from neosr_b200 import _lib
    L = _lib.lib()
    assert L.nsr_sti_bytes(131072, 180) == 1024 * 3 * 32768
    assert L.nsr_sti_bytes(100, 64) == 32768
    assert L.nsr_packed_weight_bytes(180, 180, 3, 3, 0) == L.nsr_packed_weight_bytes(180, 180, 3, 3, 1)
    assert L.nsr_loss_workspace() > 0 and L.nsr_layernorm_bwd_workspace(180) > 0


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from neosr_b200.archs import build_network
    net = build_network({"type": "swinir_small", "upscale": 2})
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.rand(1, 3, 16, 16))
    from neosr_b200.models import build_model
    with pytest.raises(RuntimeError, match="CUDA"):
        build_model({"model_type": "image", "network_g": {"type": "swinir_small"}, "is_train": True})
