"""ctypes binding of libneosr_b200.so (the C ABI declared in include/neosr_b200.h).

The library is built in-tree (`make -C neosr_b200/csrc`, or `__graft_entry__.build()`).
There is NO fallback: if the shared object is missing or a symbol is absent the import of
any compute path raises, so a silent PyTorch/eager path can never stand in for the kernels.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libneosr_b200.so"
CSRC = _PKG / "csrc"

c_float_p = C.c_void_p  # device pointers are passed as raw addresses


class NsrConv(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("pad", C.c_int32),
        ("x_ld", C.c_int32), ("y_ld", C.c_int32),
        ("act", C.c_int32), ("act_slope", C.c_float),
        ("actgrad", C.c_int32), ("actgrad_slope", C.c_float),
        ("engine", C.c_int32),
        ("x", C.c_void_p), ("w_packed", C.c_void_p), ("bias", C.c_void_p), ("prelu", C.c_void_p),
        ("aux", C.c_void_p), ("row_scale", C.c_void_p), ("residual", C.c_void_p),
        ("y_pre", C.c_void_p), ("y", C.c_void_p), ("x_sti", C.c_void_p), ("y_sti", C.c_void_p),
        ("res_ld", C.c_int32), ("aux_ld", C.c_int32), ("pre_mode", C.c_int32), ("sti_win", C.c_int32),
        ("aux_mode", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class NsrWgrad(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("pad", C.c_int32),
        ("x_ld", C.c_int32), ("dy_ld", C.c_int32), ("engine", C.c_int32),
        ("x", C.c_void_p), ("dy", C.c_void_p), ("dw", C.c_void_p), ("dbias", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("x_sti", C.c_void_p), ("dy_sti", C.c_void_p),
    ]


class NsrPackEntry(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w", "bias", "packed_fprop", "packed_dgrad", "bias_out", "row_map", "col_map")] + \
               [(n, C.c_int32) for n in ("cout", "cin", "kh", "kw", "src_cin", "reserved")] + [("block_base", C.c_int64)]


class NsrReduceEntry(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("partial", "dw", "dbias", "row_map", "col_map")] + \
               [(n, C.c_int32) for n in ("splitk", "p_rows", "p_cols", "cout", "cin", "bias_col")] + [("block_base", C.c_int64)]


class NsrAttnBiasEntry(C.Structure):
    _fields_ = [("partial", C.c_void_p), ("dbias_table", C.c_void_p)] + \
               [(n, C.c_int32) for n in ("gx", "heads", "ws", "reserved")]


class NsrParamEntry(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("p", "g", "exp_avg", "exp_avg_sq", "exp_avg_diff", "z", "neg_pre_grad", "ema")] + \
               [("n", C.c_int64), ("chunk_base", C.c_int64)]


class NsrAdanSF(C.Structure):
    _fields_ = [(n, C.c_float) for n in
                ("beta1", "one_minus_beta1", "beta2", "one_minus_beta2", "beta3", "one_minus_beta3",
                 "bias_correction3_sqrt", "eps", "decay", "ckp1", "step_size", "step_size_diff", "lr")] + \
               [("schedule_free", C.c_int32), ("first_step", C.c_int32), ("max_norm", C.c_float),
                ("ema_lerp", C.c_float), ("ema_first", C.c_int32)]


class NsrAdamW(C.Structure):
    _fields_ = [(n, C.c_float) for n in
                ("beta1", "one_minus_beta1", "beta2", "one_minus_beta2", "eps", "decay", "step_size",
                 "bias_correction2_sqrt", "max_norm", "ema_lerp")] + [("ema_first", C.c_int32)]


OPT_CHUNK = 4096
ACT = {"none": 0, "relu": 1, "lrelu": 2, "gelu": 3, "prelu": 4, "mulaux": 5}
ENGINE = {"auto": 0, "simt": 1, "tcgen05": 2, "mma_sync": 3, "bf16": 4}

_i, _f, _p, _z, _l = C.c_int, C.c_float, C.c_void_p, C.c_size_t, C.c_int64
# name -> (restype, argtypes): every symbol include/neosr_b200.h declares.
SIGNATURES = {
    "nsr_last_error": (C.c_char_p, []),
    "nsr_version": (_i, []),
    "nsr_device_supports_tcgen05": (_i, []),
    "nsr_conv_fprop": (_i, [C.POINTER(NsrConv), _p]),
    "nsr_conv_fprop_workspace": (_z, [C.POINTER(NsrConv)]),
    "nsr_packed_weight_bytes": (_z, [_i, _i, _i, _i, _i]),
    "nsr_pack_weight": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "nsr_pack_weight_pair": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "nsr_pack_entry_blocks": (_l, [_i, _i, _i, _i]),
    "nsr_pack_weights_multi": (_i, [_p, _i, _l, _p]),
    "nsr_conv_wgrad_workspace": (_z, [C.POINTER(NsrWgrad)]),
    "nsr_conv_wgrad": (_i, [C.POINTER(NsrWgrad), _p]),
    "nsr_conv_wgrad_partial_workspace": (_z, [C.POINTER(NsrWgrad)]),
    "nsr_conv_wgrad_partial": (_i, [C.POINTER(NsrWgrad), C.POINTER(C.c_int), _p]),
    "nsr_reduce_entry_blocks": (_l, [_i, _i]),
    "nsr_wgrad_finalize_multi": (_i, [_p, _i, _l, _p]),
    "nsr_nchw_to_nhwc_affine": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "nsr_nhwc_to_nchw_affine": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "nsr_pixel_shuffle_nhwc": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "nsr_maxpool2_nhwc": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "nsr_maxpool2_relu_bwd_nhwc": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "nsr_axpby": (_i, [_p, _f, _p, _f, _p, _z, _p]),
    "nsr_axpby2d": (_i, [_p, _i, _f, _p, _i, _f, _p, _i, C.c_longlong, _i, _p]),
    "nsr_actgrad_mul2d": (_i, [_p, _i, _p, _i, _p, _i, C.c_longlong, _i, _i, _f, _p]),
    "nsr_nearest_up2_nhwc": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "nsr_nearest_up2_bwd_nhwc": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "nsr_mish_fwd": (_i, [_p, _p, _z, _p]),
    "nsr_mish_bwd": (_i, [_p, _p, _p, _z, _p]),
    "nsr_mul_sigmoid_fwd": (_i, [_p, _p, _p, _z, _p]),
    "nsr_mul_sigmoid_bwd": (_i, [_p, _p, _p, _p, _p, _z, _p]),
    "nsr_add_repeat_interleave": (_i, [_p, _p, _z, _i, _i, _p]),
    "nsr_groupnorm_workspace": (_z, [_i, _i, _i]),
    "nsr_groupnorm_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p, _z, _p]),
    "nsr_groupnorm_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _z, _p]),
    "nsr_wgrad_split": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "nsr_conv_lk16_fprop": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "nsr_conv_lk16_wgrad_workspace": (_z, [_i, _i, _i, _i]),
    "nsr_conv_lk16_wgrad": (_i, [_p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _p, _z, _p]),
    "nsr_bilinear_up2_nhwc": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "nsr_bilinear_up2_bwd_nhwc": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "nsr_conv4x4s2_remap": (_i, [_p, _p, _i, _i, _i, _p]),
    "nsr_spectral_norm_workspace": (_z, [_i, _i]),
    "nsr_spectral_norm_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _f, _p, _z, _p]),
    "nsr_spectral_norm_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _z, _p]),
    "nsr_prelu_bwd_workspace": (_z, [_i]),
    "nsr_prelu_bwd": (_i, [_p, _p, _p, _p, _p, C.c_longlong, _i, _p, _z, _p]),
    "nsr_nhwc_to_nchw_add_nearest": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "nsr_actgrad_mul": (_i, [_p, _p, _p, _p, _z, _i, _f, _p]),
    "nsr_layernorm_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _f, _p, _p]),
    "nsr_sti_bytes": (_z, [C.c_longlong, _i]),
    "nsr_sti_from_f32": (_i, [_p, _i, C.c_longlong, _i, _p, _p]),
    "nsr_sti_to_f32": (_i, [_p, C.c_longlong, _i, _p, _i, _p]),
    "nsr_layernorm_bwd_workspace": (_z, [_i]),
    "nsr_layernorm_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _z, _p, _p]),
    "nsr_layernorm_bwd_blocks": (_i, [_i]),
    "nsr_layernorm_bwd2": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _z, _p, _p]),
    "nsr_window_attn_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p]),
    "nsr_window_attn_bwd_workspace": (_z, [_i, _i]),
    "nsr_window_attn_wsti_channels": (_i, [_i]),
    "nsr_window_attn_wsti_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "nsr_window_attn_wsti_bwd": (_i, [_p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _p, _z, _p]),
    "nsr_window_attn_wsti_bwd_gx": (_i, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "nsr_window_attn_dbias_multi": (_i, [_p, _i, _i, _i, _p]),
    "nsr_gather2d": (_i, [_p, _i, _p, _p, _p, _i, _i, _p]),
    "nsr_window_attn_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _p, _z, _p, _p]),
    "nsr_loss_workspace": (_z, []),
    "nsr_l1_loss": (_i, [_p, _p, _p, _z, _f, _p, _p, _p, _p]),
    "nsr_charbonnier_loss": (_i, [_p, _p, _p, _z, _f, _f, _f, _f, _p, _p, _p, _p]),
    "nsr_bce_logits_loss": (_i, [_p, _p, _z, _f, _f, _p, _p, _p, _p]),
    "nsr_grad_sumsq_workspace": (_z, []),
    "nsr_grad_sumsq": (_i, [_p, _i, _l, _p, _p, _p]),
    "nsr_adan_sf_step": (_i, [_p, _i, _l, C.POINTER(NsrAdanSF), _p, _p]),
    "nsr_adan_sf_step_dev": (_i, [_p, _i, _l, _p, _p, _p]),
    "nsr_adamw_step": (_i, [_p, _i, _l, C.POINTER(NsrAdamW), _p, _p]),
    "nsr_adamw_step_dev": (_i, [_p, _i, _l, _p, _p, _p]),
    "nsr_fsam_first_step": (_i, [_p, _i, _l, _f, _f, _f, _i, _i, _p, _p, _p]),
    "nsr_fsam_restore": (_i, [_p, _i, _l, _p]),
    "nsr_filter2d": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "nsr_resize": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _f, _f, _p]),
    "nsr_gaussian_noise": (_i, [_p, _p, _p, _p, _i, _p, _p, _i, _i, _i, C.c_uint64, _p]),
    "nsr_poisson_noise_workspace": (_z, [_i]),
    "nsr_poisson_noise": (_i, [_p, _p, _p, _p, _i, _p, _p, _i, _i, _i, C.c_uint64, _p, _z, _p]),
    "nsr_jpeg": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "nsr_crop": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "nsr_pool_swap": (_i, [_p, _p, _p, _p, _i, _z, _i, _p]),
    "nsr_resize_aa": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p]),
    "nsr_batch_mix": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _i, _i, _i, _i, _p]),
    "nsr_avgpool2": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "nsr_ssim_scale_workspace": (_z, [_i, _i, _i]),
    "nsr_ssim_scale_fwd": (_i, [_p, _p, _p, _i, _f, _f, _i, _p, _p, _i, _i, _i, _p, _z, _p]),
    "nsr_msssim_finalize": (_i, [_p, _p, _i, _f, _p, _p, _p, _p]),
    "nsr_ssim_scale_bwd": (_i, [_p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _p, _i, _i, _i, _p]),
    "nsr_clamp": (_i, [_p, _p, _z, _f, _f, _p]),
    "nsr_corr2d_zero_ext": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "nsr_consistency_workspace": (_z, [_i, _i, _i]),
    "nsr_consistency_fwd": (_i, [_p, _p, _p, _p, _f, _f, _i, _f, _p, _p, _p, _p, _i, _i, _i, _p, _z, _p]),
    "nsr_consistency_bwd": (_i, [_p, _p, _p, _p, _p, _f, _f, _p, _p, _i, _i, _i, _p, _p]),
    "nsr_reflect_fold": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "nsr_xwin_attn_stat_floats": (_z, [_i, _i, _i, _i, _i]),
    "nsr_xwin_attn_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "nsr_xwin_attn_bwd_workspace": (_z, [_i, _i, _i, _i, _i, _i, _i]),
    "nsr_xwin_attn_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _p, _z, _p]),
    "nsr_channel_mean": (_i, [_p, _p, _p, _i, _i, _i, _f, _p]),
    "nsr_channel_gate_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "nsr_channel_scale_add": (_i, [_p, _p, _p, _i, _i, _i, _f, _i, _p]),
    "nsr_channel_gate_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "nsr_channel_scale_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _f, _p]),
}

_lib = None


def build(verbose: bool = False) -> Path:
    """Compile the CUDA sources for sm_100a with nvcc (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(CSRC), "-j", str(os.cpu_count() or 4)],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"building libneosr_b200.so failed:\n{r.stdout[-4000:]}\n{r.stderr[-4000:]}")
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the shared library (once) and bind every declared symbol; raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: the neosr_b200 CUDA extension is not built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C neosr_b200/csrc`. "
            "There is no CPU/PyTorch fallback for this path.")
    dll = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(dll, name)  # AttributeError if the symbol is missing -> loud failure
        fn.restype = res
        fn.argtypes = args
    _lib = dll
    return dll


class NsrError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().nsr_last_error().decode(errors="replace")
        raise NsrError(f"{what or 'neosr_b200'} failed (code {rc}): {msg}")
