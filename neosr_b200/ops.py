"""Tensor-level wrappers over the C ABI (include/neosr_b200.h).

Everything here takes/returns CUDA fp32 torch tensors only as *memory*: the wrappers pull
`data_ptr()` and the current CUDA stream and call into libneosr_b200.so.  No torch math runs
on the hot path.  Activations are NHWC ("tokens"): a tensor of shape [B, H, W, C] (contiguous).
"""
from __future__ import annotations

import functools
import os

import ctypes as C
import math

import numpy as _np
import torch
from torch import Tensor

from . import _lib
from ._lib import ACT, ENGINE, NsrConv, NsrWgrad, check


WSTI_ATTN_ENGINE = os.environ.get("NSR_WSTI_ATTN", "auto")  # auto / tcgen05 / mma_sync kernels behind nsr_window_attn_wsti_*
WSTI_ENABLED = os.environ.get("NSR_WSTI", "1") != "0"  # window-ordered attention operands for SwinIR (A/B switch)
XWIN_TENSOR_CORES = int(os.environ.get("NSR_XWIN_TC", "3"))  # HAT window attention engine mask: bit 0 forward, bit 1 backward on mma.sync; 0 = exact fp32
BIAS_COLUMN_ENABLED = True  # bias gradients from the ones channel of split tile images (tests flip it)
LK16_ENABLED = True   # route 16->16-channel k>=7 convs to the dedicated large-kernel kernels (tests flip it)
# what engine="auto" resolves to ("auto" | "simt" | "tcgen05" | "bf16"); tests flip it.  "bf16" = the mixed-precision mode of
# `use_amp` + `bfloat16`: routed like "auto", the tcgen05 contractions issue one bf16 pass instead of three (NSR_ENGINE_BF16)
DEFAULT_ENGINE = "auto"


def _contraction_engine() -> int:
    """Engine code of the wgrad helpers that always take the tcgen05 path (split-tile-image operands)."""
    return ENGINE["bf16" if DEFAULT_ENGINE == "bf16" else "auto"]
LAUNCHES = 0          # kernels launched through this module (claim reported by bench.py)
PROFILE: list | None = None  # when a list: (kernel, shape-key, flops, bytes, start_evt, end_evt) per call


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


class _prof:
    """Context manager: CUDA events around one C-ABI call when ops.PROFILE is a list."""

    def __init__(self, name, key, flops=0.0, nbytes=0.0):
        self.on = PROFILE is not None
        if self.on:
            self.rec = [name, key, flops, nbytes, torch.cuda.Event(enable_timing=True),
                        torch.cuda.Event(enable_timing=True)]

    def __enter__(self):
        if self.on:
            self.rec[4].record()

    def __exit__(self, *a):
        if self.on:
            self.rec[5].record()
            PROFILE.append(tuple(self.rec))


def _p(t: Tensor | None):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t: Tensor | None, name: str):
    if t is None:
        return
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous CUDA fp32 tensor, got {t.dtype} {t.device} "
                         f"contiguous={t.is_contiguous()}")


class Scratch:
    """One growing device scratch buffer shared by all kernels that need workspace.
    Safe because every launch is stream-ordered on the same stream."""

    def __init__(self):
        self.buf: Tensor | None = None

    def get(self, nbytes: int, device) -> Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        return self.buf


_scratch: dict = {}


def scratch(nbytes: int, device) -> Tensor:
    if torch.cuda.is_current_stream_capturing():
        # Inside a CUDA-graph capture: memory comes from the graph's private pool and dies with the graph, and the raw
        # handle of the capture stream is recycled by later captures - a cached buffer would dangle (illegal address on
        # a later graph's replay).  Allocate per call; the pool reuses the block in capture order.
        return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    s = _scratch.get(key)
    if s is None:
        s = _scratch[key] = Scratch()
    return s.get(nbytes, device)


# ----------------------------------------------------------------------------- split tile images
class STI:
    """A token tensor [B,H,W,C] stored as a split tile image (see nsr_sti_bytes in the header):
    bf16 hi + bf16 lo in 128-row x 64-channel swizzled blocks, the layout tcgen05 contractions
    bulk-copy straight into shared memory.  Same bytes as fp32."""

    def __init__(self, shape, device):
        self.shape = tuple(shape)
        B, H, W, Cc = self.shape
        self.rows, self.c = B * H * W, Cc
        n = _lib.lib().nsr_sti_bytes(self.rows, Cc)
        # rows beyond `rows` must read as zero (they enter the wgrad reduction): zero-fill when padded
        alloc = torch.zeros if self.rows % 128 else torch.empty
        self.buf = alloc(n, dtype=torch.uint8, device=device)
        self.device = device
        self.ones = False  # set by producers that write 1.0 into the first padding channel (bias-gradient column)

    def data_ptr(self):
        return self.buf.data_ptr()

    @staticmethod
    def from_f32(x: Tensor) -> "STI":
        _chk(x, "x")
        s = STI(x.shape, x.device)
        check(_lib.lib().nsr_sti_from_f32(x.data_ptr(), s.c, s.rows, s.c, s.data_ptr(), _stream()), "nsr_sti_from_f32")
        _count(1)
        return s

    def to_f32(self) -> Tensor:
        y = torch.empty(self.shape, dtype=torch.float32, device=self.device)
        check(_lib.lib().nsr_sti_to_f32(self.data_ptr(), self.rows, self.c, y.data_ptr(), self.c, _stream()), "nsr_sti_to_f32")
        _count(1)
        return y


class Slab:
    """A channel slice [c0, c0 + c) of a wider NHWC buffer [B,H,W,LD] (ESRGAN dense blocks read and
    write growing channel slabs in place instead of torch.cat copies, esrgan_arch.py:109-116)."""

    def __init__(self, base: Tensor, c0: int, c: int):
        _chk(base, "slab base")
        B, H, W, LD = base.shape
        if c0 < 0 or c0 + c > LD or c0 % 4 or LD % 4:
            raise ValueError("Slab: channel range must lie in the buffer and be 16-byte aligned")
        self.base, self.c0, self.c, self.ld = base, c0, c, LD
        self.shape = (B, H, W, c)
        self.device = base.device

    def data_ptr(self):
        return self.base.data_ptr() + 4 * self.c0

    def numel(self):
        return self.shape[0] * self.shape[1] * self.shape[2] * self.c


def _ptr_ld(t):
    """(device pointer, leading dim) of a contiguous NHWC tensor or a Slab."""
    if t is None:
        return None, 0
    if isinstance(t, Slab):
        return t.data_ptr(), t.ld
    _chk(t, "tensor")
    return t.data_ptr(), t.shape[-1]


def sti_enabled() -> bool:
    """Split-tile-image operands need the tcgen05 engine (sm_100) and are skipped when a test
    forces the exact-fp32 engine."""
    return DEFAULT_ENGINE != "simt" and bool(_lib.lib().nsr_device_supports_tcgen05())


# ----------------------------------------------------------------------------- weights
class PackedWeight:
    """fprop- and dgrad-flavoured packed copies of one Conv2d/Linear weight
    (reference layout [cout, cin, kh, kw] or [cout, cin])."""

    def __init__(self, weight: Tensor, need_dgrad: bool = True):
        self.weight = weight
        self.cout, self.cin = weight.shape[0], weight.shape[1]
        self.kh = weight.shape[2] if weight.dim() == 4 else 1
        self.kw = weight.shape[3] if weight.dim() == 4 else 1
        L = _lib.lib()
        dev = weight.device
        self.fprop = torch.empty(L.nsr_packed_weight_bytes(self.cout, self.cin, self.kh, self.kw, 0),
                                 dtype=torch.uint8, device=dev)
        self.dgrad = torch.empty(L.nsr_packed_weight_bytes(self.cout, self.cin, self.kh, self.kw, 1),
                                 dtype=torch.uint8, device=dev) if need_dgrad else None
        self._version = None

    def version_key(self):
        w = self.weight
        return (w.data_ptr(), w._version)

    def stale(self) -> bool:
        return self.version_key() != self._version

    def pack_entry(self) -> dict:
        """Fields of this weight's NsrPackEntry (nsr_pack_weights_multi)."""
        w = self.weight
        return dict(w=w.data_ptr(), bias=None, packed_fprop=self.fprop.data_ptr(), packed_dgrad=_p(self.dgrad), bias_out=None,
                    row_map=None, col_map=None, cout=self.cout, cin=self.cin, kh=self.kh, kw=self.kw, src_cin=self.cin)

    def refresh(self, force: bool = False) -> "PackedWeight":
        """Re-pack if the parameter changed since the last pack (tracked by tensor version)."""
        w = self.weight
        ver = (w.data_ptr(), w._version)
        if not force and ver == self._version:
            return self
        _chk(w, "weight")
        L = _lib.lib()
        st = _stream()
        check(L.nsr_pack_weight_pair(w.data_ptr(), self.cout, self.cin, self.kh, self.kw, self.fprop.data_ptr(),
                                     _p(self.dgrad), st), "nsr_pack_weight_pair")
        _count(1)
        self._version = ver
        return self


class MappedPackedWeight(PackedWeight):
    """Packed copy of a Linear weight [cout, cin] whose output rows and / or input columns are re-indexed through int
    maps (-1 = a zero row / column), plus the matching padded bias.  Used for the head-padded qkv (rows) and proj
    (columns) weights behind the window-ordered attention operands: re-gathered (`nsr_gather2d`) and re-packed whenever
    the parameter's version changes, like any PackedWeight."""

    def __init__(self, weight: Tensor, bias: Tensor | None, row_map=None, col_map=None, need_dgrad: bool = True):
        dev = weight.device
        rows = len(row_map) if row_map is not None else weight.shape[0]
        cols = len(col_map) if col_map is not None else weight.shape[1]
        self.src, self.src_bias = weight, bias
        self.row_map = torch.tensor(row_map, dtype=torch.int32, device=dev) if row_map is not None else None
        self.col_map = torch.tensor(col_map, dtype=torch.int32, device=dev) if col_map is not None else None
        self.padded = torch.zeros((rows, cols), dtype=torch.float32, device=dev)
        self.bias_padded = torch.zeros(rows, dtype=torch.float32, device=dev) if bias is not None else None
        super().__init__(self.padded, need_dgrad=need_dgrad)
        self.weight = weight  # identity of the source parameter (ParamSet.pw compares data pointers)

    def version_key(self):
        w, b = self.src, self.src_bias
        return (w.data_ptr(), w._version, None if b is None else b._version)

    def pack_entry(self) -> dict:
        w, b = self.src, self.src_bias
        return dict(w=w.data_ptr(), bias=_p(b), packed_fprop=self.fprop.data_ptr(), packed_dgrad=_p(self.dgrad),
                    bias_out=_p(self.bias_padded), row_map=_p(self.row_map), col_map=_p(self.col_map), cout=self.cout,
                    cin=self.cin, kh=1, kw=1, src_cin=w.shape[1])

    def refresh(self, force: bool = False) -> "MappedPackedWeight":
        w, b = self.src, self.src_bias
        ver = self.version_key()
        if not force and ver == self._version:
            return self
        _chk(w, "weight")
        L = _lib.lib()
        st = _stream()
        check(L.nsr_gather2d(w.data_ptr(), w.shape[1], _p(self.row_map), _p(self.col_map), self.padded.data_ptr(),
                             self.padded.shape[0], self.padded.shape[1], st), "nsr_gather2d")
        if b is not None:
            check(L.nsr_gather2d(b.data_ptr(), b.shape[0], None, _p(self.row_map), self.bias_padded.data_ptr(), 1,
                                 self.bias_padded.shape[0], st), "nsr_gather2d")
        check(L.nsr_pack_weight_pair(self.padded.data_ptr(), self.cout, self.cin, 1, 1, self.fprop.data_ptr(),
                                     _p(self.dgrad), st), "nsr_pack_weight_pair")
        _count(2 if b is None else 3)
        self._version = ver
        return self


def pack_weights_multi(pks: list, cache: dict) -> None:
    """One launch for a list of PackedWeight / MappedPackedWeight objects (engine.ParamSet.pack_all); the device table is
    cached per set of objects (their buffers never move)."""
    key = tuple(id(pk) for pk in pks)
    ent = cache.get(key)
    L = _lib.lib()
    if ent is None:
        arr = (_lib.NsrPackEntry * len(pks))()
        base = 0
        for i, pk in enumerate(pks):
            _chk(pk.weight, "weight")
            for k, v in pk.pack_entry().items():
                setattr(arr[i], k, v)
            arr[i].block_base = base
            base += L.nsr_pack_entry_blocks(pk.cout, pk.cin, pk.kh, pk.kw)
        dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(pks[0].weight.device)
        ent = cache[key] = (dev, len(pks), base)
        if len(cache) > 8:  # stale object sets (re-built ParamSets)
            cache.pop(next(iter(cache)))
    dev, n, blocks = ent
    check(L.nsr_pack_weights_multi(dev.data_ptr(), n, blocks, _stream()), "nsr_pack_weights_multi")
    _count(1)
    for pk in pks:
        pk._version = pk.version_key()


@functools.lru_cache(maxsize=64)
def head_pad_map(c: int, heads: int, groups: int) -> tuple:
    """Channel map of the window-ordered attention operands: `groups` blocks (q | k | v, or 1) of G =
    nsr_window_attn_wsti_channels(heads) channels, head h of a group at [h*32, h*32 + c/heads), the rest -1 (zero)."""
    G = _lib.lib().nsr_window_attn_wsti_channels(heads)
    d = c // heads
    m = [-1] * (groups * G)
    for s in range(groups):
        for h in range(heads):
            for i in range(d):
                m[s * G + h * 32 + i] = s * c + h * d + i
    return tuple(m)


# fc1 hands gelu'(pre) to fc2's dgrad as 16-bit codes (NsrConv.aux_mode) on the split-tile-image path; NSR_AGC_U16=0: fp32
AGC_U16 = os.environ.get("NSR_AGC_U16", "1") != "0"
# the LayerNorm backward kernels of a Swin block pass the token-stream gradient on as a tile image only; NSR_LN_STI_RES=0: fp32 + image
LN_STI_RES = os.environ.get("NSR_LN_STI_RES", "1") != "0"


def wsti_supported(c: int, heads: int, ws: int) -> bool:
    """Window-ordered attention operands: tcgen05 engine + the mma attention kernels' shape limits."""
    return sti_enabled() and _attn_mma_ok(c, heads, ws)


# ----------------------------------------------------------------------------- contraction
def conv_fprop(x: Tensor, pw: PackedWeight, bias: Tensor | None = None, *, dgrad: bool = False,
               act: str = "none", act_slope: float = 0.0, actgrad: str = "none", actgrad_slope: float = 0.0,
               aux: Tensor | None = None, prelu: Tensor | None = None, row_scale: Tensor | None = None,
               residual: Tensor | None = None, want_pre: bool = False, out: Tensor | None = None,
               engine: str = "auto", sti_out: bool = False, f32_out: bool = True, pre_is_actgrad: bool = False,
               sti_win: tuple | None = None, pre_u16: bool = False):
    """y = epilogue(conv(x, w)); x is [B,H,W,Cin] NHWC fp32 or an STI (1x1 only).  With dgrad=True
    the dgrad-packed filter is used and the roles of cin/cout swap (x is then dY [B,H,W,Cout]).
    sti_out / f32_out select the output formats: returns y (fp32), or the STI, or (y, sti).
    sti_win = (ws, shift): the STI's rows are written in window order (NsrConv.sti_win).
    pre_u16 (with pre_is_actgrad, STI-only output): y_pre is returned as int16 activation-gradient codes (NsrConv.aux_mode);
    an int16 `aux` is read as such codes."""
    x_is_sti = isinstance(x, STI)
    if not x_is_sti and not isinstance(x, Slab):
        _chk(x, "x")
    B, H, W, cx = x.shape
    cin, cout = (pw.cout, pw.cin) if dgrad else (pw.cin, pw.cout)
    if cx != cin:
        raise ValueError(f"conv_fprop: x has {cx} channels, weight expects {cin}")
    for t, n in ((bias, "bias"), (prelu, "prelu"), (row_scale, "row_scale")):
        _chk(t, n)
    x_ptr, x_ld = (None, cin) if x_is_sti else _ptr_ld(x)
    if (LK16_ENABLED and cin == 16 and cout == 16 and pw.kh == pw.kw and pw.kh >= 7 and not x_is_sti and act == "none"
            and actgrad == "none" and aux is None and prelu is None and row_scale is None and residual is None
            and not want_pre and not sti_out and f32_out and engine == "auto"):
        # RealPLKSR's 16-channel large-kernel conv: dedicated exact-fp32 kernel (csrc/conv_lk.cu)
        y = out if out is not None else torch.empty((B, H, W, cout), dtype=torch.float32, device=x.device)
        y_ptr, y_ld = _ptr_ld(y)
        wbuf = pw.dgrad if dgrad else pw.fprop
        with _prof("conv_lk16_" + ("dgrad" if dgrad else "fprop"), (B * H * W, cin, cout, pw.kh),
                   2.0 * B * H * W * cin * cout * pw.kh * pw.kw, 4.0 * B * H * W * (cin + cout)):
            check(_lib.lib().nsr_conv_lk16_fprop(x_ptr, x_ld, wbuf.data_ptr(), _p(bias), y_ptr, y_ld, B, H, W, pw.kh, _stream()),
                  "nsr_conv_lk16_fprop")
        _count(1)
        return y
    if aux is not None and not isinstance(aux, Slab) and aux.dtype == torch.int16:  # activation-gradient codes
        if not (aux.is_cuda and aux.is_contiguous()):
            raise ValueError("aux: expected a contiguous CUDA int16 tensor of activation-gradient codes")
        aux_ptr, aux_ld = aux.data_ptr(), aux.shape[-1]
    else:
        aux_ptr, aux_ld = _ptr_ld(aux)
    res_ptr, res_ld = _ptr_ld(residual)
    y = None
    if f32_out:
        y = out if out is not None else torch.empty((B, H, W, cout), dtype=torch.float32, device=x.device)
    y_ptr, y_ld = _ptr_ld(y) if y is not None else (None, cout)
    if want_pre and isinstance(y, Slab):
        raise ValueError("conv_fprop: y_pre shares y's leading dim; write y to a dense tensor when want_pre")
    y_sti = STI((B, H, W, cout), x.device) if sti_out else None
    if y_sti is not None:
        y_sti.ones = cout % 64 != 0
    if y is None and y_sti is None:
        raise ValueError("conv_fprop: no output format selected")
    if pre_u16 and not (want_pre and pre_is_actgrad and sti_out and not f32_out):
        raise ValueError("conv_fprop: pre_u16 needs want_pre, pre_is_actgrad and a split-tile-image-only output")
    y_pre = torch.empty((B, H, W, cout), dtype=torch.int16 if pre_u16 else torch.float32, device=x.device) if want_pre else None
    aux_u16 = aux is not None and not isinstance(aux, Slab) and aux.dtype == torch.int16
    d = NsrConv(batch=B, h=H, w=W, cin=cin, cout=cout, kh=pw.kh, kw=pw.kw, pad=pw.kh // 2,
                x_ld=x_ld, y_ld=y_ld, act=ACT[act], act_slope=act_slope, actgrad=ACT[actgrad],
                actgrad_slope=actgrad_slope, engine=ENGINE[DEFAULT_ENGINE if engine == "auto" else engine],
                x=x_ptr, w_packed=(pw.dgrad if dgrad else pw.fprop).data_ptr(),
                bias=_p(bias), prelu=_p(prelu), aux=aux_ptr, row_scale=_p(row_scale), residual=res_ptr,
                y_pre=_p(y_pre), y=y_ptr, x_sti=x.data_ptr() if x_is_sti else None, y_sti=_p(y_sti),
                res_ld=res_ld if res_ld != y_ld else 0, aux_ld=aux_ld if aux_ld != y_ld else 0,
                pre_mode=(2 if pre_u16 else 1) if pre_is_actgrad else 0, aux_mode=2 if aux_u16 else 0,
                sti_win=(sti_win[0] | (sti_win[1] << 16)) if sti_win else 0,
                workspace=None, workspace_bytes=0)
    if min(cin, cout) <= 4:  # image-side convs: im2col + tensor-core contraction needs scratch
        need = _lib.lib().nsr_conv_fprop_workspace(C.byref(d))
        if need:
            ws = scratch(need, x.device)
            d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    M = B * H * W
    with _prof(("conv_dgrad" if dgrad else "conv_fprop") + ("_sti" if x_is_sti else ""), (M, cin, cout, pw.kh),
               2.0 * M * cin * cout * pw.kh * pw.kw, 4.0 * M * (cin + cout)):
        check(_lib.lib().nsr_conv_fprop(C.byref(d), _stream()), "nsr_conv_fprop")
    _count(1)
    res = y if y_sti is None else (y_sti if y is None else (y, y_sti))
    return (res, y_pre) if want_pre else res


def conv_wgrad(x, dy, dw: Tensor, dbias: Tensor | None, kh: int, kw: int, engine: str = "auto",
               x_sti: STI | None = None, dy_sti: STI | None = None):
    """dw[cout,cin,kh,kw] (+ dbias) from x [B,H,W,Cin] and dy [B,H,W,Cout]; overwrites dw/dbias.
    x / dy may be None when their split tile images are given (1x1 contractions)."""
    _chk(dw, "dw"), _chk(dbias, "dbias")
    xs, ds = (x if x is not None else x_sti), (dy if dy is not None else dy_sti)
    B, H, W, cin = xs.shape
    cout = ds.shape[-1]
    x, dy = (x if x is not None else xs), (dy if dy is not None else ds)
    if dy.shape[:3] != x.shape[:3] or dw.numel() != cout * cin * kh * kw:
        raise ValueError(f"conv_wgrad: shape mismatch x{tuple(x.shape)} dy{tuple(dy.shape)} dw{tuple(dw.shape)}")
    x_ptr, x_ld = (None, cin) if isinstance(x, STI) else _ptr_ld(x)
    dy_ptr, dy_ld = (None, cout) if isinstance(dy, STI) else _ptr_ld(dy)
    if (LK16_ENABLED and cin == 16 and cout == 16 and kh == kw and kh >= 7 and engine == "auto" and x_ptr is not None
            and dy_ptr is not None):
        L = _lib.lib()
        ws = scratch(L.nsr_conv_lk16_wgrad_workspace(B, H, W, kh), dw.device)
        with _prof("conv_lk16_wgrad", (B * H * W, cin, cout, kh), 2.0 * B * H * W * cin * cout * kh * kw,
                   4.0 * B * H * W * (cin + cout)):
            check(L.nsr_conv_lk16_wgrad(x_ptr, x_ld, dy_ptr, dy_ld, dw.data_ptr(), _p(dbias), B, H, W, kh, ws.data_ptr(),
                                        ws.numel(), _stream()), "nsr_conv_lk16_wgrad")
        _count(2)
        return
    # bias gradient for free: the x image carries 1.0 in channel cin, so dW over cin + 4 channels has dbias as column cin
    fused_bias = (BIAS_COLUMN_ENABLED and dbias is not None and x_sti is not None and dy_sti is not None and x_sti.ones
                  and kh == 1 and kw == 1 and cin % 64 != 0 and cin + 4 <= (cin + 63) // 64 * 64)
    tmp = None
    if fused_bias:
        tmp = torch.empty((cout, cin + 4), dtype=torch.float32, device=dw.device)
    d = NsrWgrad(batch=B, h=H, w=W, cin=cin + 4 if fused_bias else cin, cout=cout, kh=kh, kw=kw, pad=kh // 2,
                 x_ld=x_ld + 4 if fused_bias else x_ld, dy_ld=dy_ld,
                 engine=ENGINE[DEFAULT_ENGINE if engine == "auto" else engine],
                 x=None if fused_bias else x_ptr, dy=None if fused_bias else dy_ptr,
                 dw=tmp.data_ptr() if fused_bias else dw.data_ptr(), dbias=None if fused_bias else _p(dbias),
                 workspace=None, workspace_bytes=0, x_sti=_p(x_sti), dy_sti=_p(dy_sti))
    L = _lib.lib()
    need = L.nsr_conv_wgrad_workspace(C.byref(d))
    ws = scratch(need, x.device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    M = B * H * W
    with _prof("conv_wgrad" + ("_sti" if x_sti is not None and dy_sti is not None else ""), (M, cin, cout, kh),
               2.0 * M * cin * cout * kh * kw, 4.0 * M * (cin + cout)):
        check(L.nsr_conv_wgrad(C.byref(d), _stream()), "nsr_conv_wgrad")
        if fused_bias:
            check(L.nsr_wgrad_split(tmp.data_ptr(), dw.data_ptr(), dbias.data_ptr(), cout, cin, cin + 4, _stream()), "nsr_wgrad_split")
    _count(3 if fused_bias else (4 if dbias is not None else 2))


def conv_wgrad_mapped(x_sti: STI, dy_sti: STI, dw: Tensor, dbias: Tensor | None, col_map: Tensor, ones_col: int | None):
    """Weight gradient of a Linear whose INPUT image is head-padded (x_sti: [.., G] channels, `col_map[k]` = the real
    input channel of padded channel k or -1): dW over the padded channels on the tcgen05 wgrad kernel, then un-padded
    into dw [cout, cin]; dbias = the column of the channel that carries 1.0 (`ones_col`)."""
    _chk(dw, "dw"), _chk(dbias, "dbias")
    B, H, W, G = x_sti.shape
    cout, cin = dw.shape[0], dw.shape[1]
    tmp = torch.empty((cout, G), dtype=torch.float32, device=dw.device)
    d = NsrWgrad(batch=B, h=H, w=W, cin=G, cout=cout, kh=1, kw=1, pad=0, x_ld=G, dy_ld=cout, engine=_contraction_engine(), x=None,
                 dy=None, dw=tmp.data_ptr(), dbias=None, workspace=None, workspace_bytes=0, x_sti=x_sti.data_ptr(),
                 dy_sti=dy_sti.data_ptr())
    L = _lib.lib()
    ws = scratch(L.nsr_conv_wgrad_workspace(C.byref(d)), dw.device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    M = B * H * W
    with _prof("conv_wgrad_sti", (M, G, cout, 1), 2.0 * M * G * cout, 4.0 * M * (G + cout)):
        check(L.nsr_conv_wgrad(C.byref(d), _stream()), "nsr_conv_wgrad")
        inv = _inverse_map(col_map, cin)
        check(L.nsr_gather2d(tmp.data_ptr(), G, None, inv.data_ptr(), dw.data_ptr(), cout, cin, _stream()), "nsr_gather2d")
        if dbias is not None:
            if ones_col is None:
                raise ValueError("conv_wgrad_mapped: the image carries no ones column for the bias gradient")
            oc = _const_i32((ones_col,), dw.device)
            check(L.nsr_gather2d(tmp.data_ptr(), G, None, oc.data_ptr(), dbias.data_ptr(), cout, 1, _stream()), "nsr_gather2d")
    _count(4 if dbias is not None else 3)


def conv_wgrad_mapped_rows(x_sti: STI, dy_sti: STI, dw: Tensor, dbias: Tensor | None, row_map: Tensor):
    """Weight gradient of a Linear whose OUTPUT-gradient image is head-padded (dy_sti: [.., Gout] channels, `row_map[n']` =
    the real output channel of padded channel n' or -1): dW over the padded rows on the tcgen05 wgrad kernel (bias gradient
    from the ones column of x_sti), then the real rows are gathered into dw [cout, cin] / dbias [cout]."""
    _chk(dw, "dw"), _chk(dbias, "dbias")
    B, H, W, cin = x_sti.shape
    gout = dy_sti.shape[-1]
    cout = dw.shape[0]
    fused = dbias is not None and x_sti.ones and cin % 64 != 0 and cin + 4 <= (cin + 63) // 64 * 64
    if dbias is not None and not fused:
        raise ValueError("conv_wgrad_mapped_rows: the input image carries no ones column for the bias gradient")
    cw = cin + 4 if fused else cin
    tmp = torch.empty((gout, cw), dtype=torch.float32, device=dw.device)
    d = NsrWgrad(batch=B, h=H, w=W, cin=cw, cout=gout, kh=1, kw=1, pad=0, x_ld=cw, dy_ld=gout, engine=_contraction_engine(), x=None,
                 dy=None, dw=tmp.data_ptr(), dbias=None, workspace=None, workspace_bytes=0, x_sti=x_sti.data_ptr(),
                 dy_sti=dy_sti.data_ptr())
    L = _lib.lib()
    ws = scratch(L.nsr_conv_wgrad_workspace(C.byref(d)), dw.device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    M = B * H * W
    with _prof("conv_wgrad_sti", (M, cin, gout, 1), 2.0 * M * cin * gout, 4.0 * M * (cin + gout)):
        check(L.nsr_conv_wgrad(C.byref(d), _stream()), "nsr_conv_wgrad")
        inv = _inverse_map(row_map, cout)
        check(L.nsr_gather2d(tmp.data_ptr(), cw, inv.data_ptr(), None, dw.data_ptr(), cout, cin, _stream()), "nsr_gather2d")
        if dbias is not None:
            oc = _const_i32((cin,), dw.device)
            check(L.nsr_gather2d(tmp.data_ptr(), cw, inv.data_ptr(), oc.data_ptr(), dbias.data_ptr(), cout, 1, _stream()),
                  "nsr_gather2d")
    _count(4 if dbias is not None else 3)


class DeferredWgrads:
    """1x1 weight gradients whose split-K partials are reduced LATER, all together (`nsr_wgrad_finalize_multi`): `add`
    runs only the tcgen05 contraction into a per-layer buffer, `finalize` (end of the backward pass) is one launch that
    writes every dW / dbias, un-padding head-padded rows / columns and splitting off the bias-gradient column."""

    def __init__(self):
        self.jobs: list = []
        self.attn_jobs: list = []  # window-attention bias-table gradients (nsr_window_attn_dbias_multi)
        self.buffers: dict = {}
        self.tables: dict = {}

    def add(self, key, x_sti: STI, dy_sti: STI, dw: Tensor, dbias: Tensor | None, *, row_map: Tensor | None = None,
            col_map: Tensor | None = None, bias_col: int | None = None) -> None:
        """row_map / col_map: padded -> real channel maps (-1 = padding) of dy_sti's / x_sti's channels, or None."""
        _chk(dw, "dw"), _chk(dbias, "dbias")
        B, H, W, cx = x_sti.shape
        gout = dy_sti.shape[-1]
        cout, cin = dw.shape[0], dw.shape[1]
        cw = cx
        if dbias is not None and bias_col is None:  # LayerNorm-style image: 1.0 in the first padding channel
            if not (x_sti.ones and cx % 64 != 0 and cx + 4 <= (cx + 63) // 64 * 64):
                raise ValueError("DeferredWgrads: the input image carries no ones column for the bias gradient")
            cw, bias_col = cx + 4, cx
        d = NsrWgrad(batch=B, h=H, w=W, cin=cw, cout=gout, kh=1, kw=1, pad=0, x_ld=cw, dy_ld=gout, engine=_contraction_engine(),
                     x=None, dy=None, dw=None, dbias=None, workspace=None, workspace_bytes=0, x_sti=x_sti.data_ptr(),
                     dy_sti=dy_sti.data_ptr())
        L = _lib.lib()
        need = L.nsr_conv_wgrad_partial_workspace(C.byref(d))
        if need == 0:
            raise RuntimeError("DeferredWgrads: shape not supported by the tcgen05 wgrad kernel")
        buf = self.buffers.get(key)
        if buf is None or buf.numel() < need:
            buf = self.buffers[key] = torch.empty(need, dtype=torch.uint8, device=dw.device)
        d.workspace, d.workspace_bytes = buf.data_ptr(), buf.numel()
        splitk = C.c_int(0)
        M = B * H * W
        with _prof("conv_wgrad_sti", (M, cx, gout, 1), 2.0 * M * cx * gout, 4.0 * M * (cx + gout)):
            check(L.nsr_conv_wgrad_partial(C.byref(d), C.byref(splitk), _stream()), "nsr_conv_wgrad_partial")
        _count(1)
        self.jobs.append(dict(partial=buf.data_ptr(), dw=dw.data_ptr(), dbias=_p(dbias),
                              row_map=None if row_map is None else _inverse_map(row_map, cout).data_ptr(),
                              col_map=None if col_map is None else _inverse_map(col_map, cin).data_ptr(),
                              splitk=splitk.value, p_rows=gout, p_cols=cw, cout=cout, cin=cin,
                              bias_col=-1 if bias_col is None else bias_col))

    def _finalize_attn(self) -> None:
        jobs, self.attn_jobs = self.attn_jobs, []
        key = ("attn",) + tuple(tuple(j.values()) for j in jobs)
        ent = self.tables.get(key)
        if ent is None:
            arr = (_lib.NsrAttnBiasEntry * len(jobs))()
            for i, j in enumerate(jobs):
                for k, v in j.items():
                    setattr(arr[i], k, v)
            dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(torch.device("cuda", torch.cuda.current_device()))
            ent = self.tables[key] = (dev, len(jobs), max(j["heads"] for j in jobs), max(j["ws"] for j in jobs))
            if len(self.tables) > 8:
                self.tables.pop(next(iter(self.tables)))
        dev, n, mh, mw = ent
        with _prof("nsr_window_attn_dbias_multi", (n,), 0.0, 0.0):
            check(_lib.lib().nsr_window_attn_dbias_multi(dev.data_ptr(), n, mh, mw, _stream()), "nsr_window_attn_dbias_multi")
        _count(2)

    def finalize(self) -> None:
        if self.attn_jobs:
            self._finalize_attn()
        if not self.jobs:
            return
        jobs, self.jobs = self.jobs, []
        key = tuple(tuple(j.values()) for j in jobs)
        ent = self.tables.get(key)
        L = _lib.lib()
        if ent is None:
            arr = (_lib.NsrReduceEntry * len(jobs))()
            base = 0
            for i, j in enumerate(jobs):
                for k, v in j.items():
                    setattr(arr[i], k, v)
                arr[i].block_base = base
                base += L.nsr_reduce_entry_blocks(j["cout"], j["cin"])
            dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(torch.device("cuda", torch.cuda.current_device()))
            ent = self.tables[key] = (dev, len(jobs), base)
            if len(self.tables) > 8:
                self.tables.pop(next(iter(self.tables)))
        dev, n, blocks = ent
        with _prof("nsr_wgrad_finalize_multi", (n,), 0.0, 0.0):
            check(L.nsr_wgrad_finalize_multi(dev.data_ptr(), n, blocks, _stream()), "nsr_wgrad_finalize_multi")
        _count(1)


DEFER_WGRAD = os.environ.get("NSR_DEFER_WGRAD", "1") != "0"  # A/B switch for the deferred, batched wgrad reduction
_I32_CACHE: dict = {}


def _const_i32(vals: tuple, device) -> Tensor:
    key = (vals, str(device))
    t = _I32_CACHE.get(key)
    if t is None:
        t = _I32_CACHE[key] = torch.tensor(vals, dtype=torch.int32, device=device)
    return t


def _inverse_map(col_map: Tensor, n: int) -> Tensor:
    """padded index of every real channel (col_map: padded -> real or -1), cached per map tensor."""
    key = ("inv", col_map.data_ptr(), n)
    t = _I32_CACHE.get(key)
    if t is None:
        m = col_map.tolist()
        inv = [0] * n
        for k, c in enumerate(m):
            if c >= 0:
                inv[c] = k
        t = _I32_CACHE[key] = torch.tensor(inv, dtype=torch.int32, device=col_map.device)
    return t


# ----------------------------------------------------------------------------- layout
def nchw_to_nhwc_affine(x: Tensor, scale: Tensor | None, shift: Tensor | None) -> Tensor:
    _chk(x, "x")
    B, Cc, H, W = x.shape
    y = torch.empty((B, H, W, Cc), dtype=torch.float32, device=x.device)
    with _prof("nsr_nchw_to_nhwc_affine", (x.numel(),), 0.0, 8.0 * x.numel()):
        check(_lib.lib().nsr_nchw_to_nhwc_affine(x.data_ptr(), y.data_ptr(), B, Cc, H, W, _p(scale), _p(shift), _stream()),
              "nsr_nchw_to_nhwc_affine")
    _count(1)
    return y


def nhwc_to_nchw_affine(x: Tensor, scale: Tensor | None, shift: Tensor | None) -> Tensor:
    _chk(x, "x")
    B, H, W, Cc = x.shape
    y = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    with _prof("nsr_nhwc_to_nchw_affine", (x.numel(),), 0.0, 8.0 * x.numel()):
        check(_lib.lib().nsr_nhwc_to_nchw_affine(x.data_ptr(), y.data_ptr(), B, Cc, H, W, _p(scale), _p(shift), _stream()),
              "nsr_nhwc_to_nchw_affine")
    _count(1)
    return y


def pixel_shuffle(x: Tensor, r: int) -> Tensor:
    """[B,H,W,C*r*r] -> [B,H*r,W*r,C]."""
    _chk(x, "x")
    B, H, W, Ci = x.shape
    Co = Ci // (r * r)
    y = torch.empty((B, H * r, W * r, Co), dtype=torch.float32, device=x.device)
    with _prof("nsr_pixel_shuffle_nhwc", (x.numel(),), 0.0, 8.0 * x.numel()):
        check(_lib.lib().nsr_pixel_shuffle_nhwc(x.data_ptr(), y.data_ptr(), B, H, W, Co, r, 0, _stream()),
              "nsr_pixel_shuffle_nhwc")
    _count(1)
    return y


def pixel_unshuffle(dy: Tensor, r: int) -> Tensor:
    """[B,H*r,W*r,C] -> [B,H,W,C*r*r] (autograd of pixel_shuffle)."""
    _chk(dy, "dy")
    B, Hr, Wr, Co = dy.shape
    H, W = Hr // r, Wr // r
    y = torch.empty((B, H, W, Co * r * r), dtype=torch.float32, device=dy.device)
    with _prof("nsr_pixel_unshuffle_nhwc", (dy.numel(),), 0.0, 8.0 * dy.numel()):
        check(_lib.lib().nsr_pixel_shuffle_nhwc(dy.data_ptr(), y.data_ptr(), B, H, W, Co, r, 1, _stream()),
              "nsr_pixel_shuffle_nhwc")
    _count(1)
    return y


def maxpool2(x: Tensor) -> Tensor:
    _chk(x, "x")
    B, H, W, Cc = x.shape
    y = torch.empty((B, H // 2, W // 2, Cc), dtype=torch.float32, device=x.device)
    with _prof("nsr_maxpool2_nhwc", (x.numel(),), 0.0, 5.0 * x.numel()):
        check(_lib.lib().nsr_maxpool2_nhwc(x.data_ptr(), y.data_ptr(), B, H, W, Cc, _stream()), "nsr_maxpool2_nhwc")
    _count(1)
    return y


def maxpool2_relu_bwd(x: Tensor, dy: Tensor, dextra: Tensor | None = None) -> Tensor:
    _chk(x, "x"), _chk(dy, "dy"), _chk(dextra, "dextra")
    B, H, W, Cc = x.shape
    dx = torch.empty_like(x)
    with _prof("nsr_maxpool2_relu_bwd_nhwc", (x.numel(),), 0.0, 9.0 * x.numel()):
        check(_lib.lib().nsr_maxpool2_relu_bwd_nhwc(x.data_ptr(), dy.data_ptr(), _p(dextra), dx.data_ptr(), B, H, W, Cc,
                                                    _stream()), "nsr_maxpool2_relu_bwd_nhwc")
    _count(1)
    return dx


def axpby(a: Tensor, alpha: float, b: Tensor | None, beta: float, out: Tensor | None = None) -> Tensor:
    _chk(a, "a"), _chk(b, "b")
    y = out if out is not None else torch.empty_like(a)
    with _prof("nsr_axpby", (a.numel(),), 0.0, 12.0 * a.numel()):
        check(_lib.lib().nsr_axpby(a.data_ptr(), alpha, _p(b), beta, y.data_ptr(), a.numel(), _stream()), "nsr_axpby")
    _count(1)
    return y


def axpby2d(a, alpha: float, b, beta: float, out=None):
    """out = a * alpha + b * beta over [rows, cols] views (tensors or Slabs); returns `out` (dense if not given)."""
    B, H, W, cols = a.shape
    rows = B * H * W
    y = out if out is not None else torch.empty((B, H, W, cols), dtype=torch.float32, device=a.device)
    (ap, lda), (bp, ldb), (yp, ldy) = _ptr_ld(a), _ptr_ld(b), _ptr_ld(y)
    with _prof("nsr_axpby2d", (rows, cols), 0.0, 12.0 * rows * cols):
        check(_lib.lib().nsr_axpby2d(ap, lda, alpha, bp, ldb, beta, yp, ldy, rows, cols, _stream()), "nsr_axpby2d")
    _count(1)
    return y


def actgrad_mul2d(dy, aux, act: str, slope: float = 0.0) -> Tensor:
    """dense dx = dy * act'(aux) for Slab / tensor views."""
    B, H, W, cols = dy.shape
    rows = B * H * W
    dx = torch.empty((B, H, W, cols), dtype=torch.float32, device=dy.device)
    (dp, ldd), (xp, lda) = _ptr_ld(dy), _ptr_ld(aux)
    with _prof("nsr_actgrad_mul2d", (rows, cols), 0.0, 12.0 * rows * cols):
        check(_lib.lib().nsr_actgrad_mul2d(dp, ldd, xp, lda, dx.data_ptr(), cols, rows, cols, ACT[act], slope, _stream()),
              "nsr_actgrad_mul2d")
    _count(1)
    return dx


def nearest_up2(x: Tensor) -> Tensor:
    _chk(x, "x")
    B, H, W, Cc = x.shape
    y = torch.empty((B, 2 * H, 2 * W, Cc), dtype=torch.float32, device=x.device)
    with _prof("nsr_nearest_up2_nhwc", (x.numel(),), 0.0, 20.0 * x.numel()):
        check(_lib.lib().nsr_nearest_up2_nhwc(x.data_ptr(), y.data_ptr(), B, H, W, Cc, _stream()), "nsr_nearest_up2_nhwc")
    _count(1)
    return y


def nearest_up2_bwd(dy: Tensor) -> Tensor:
    _chk(dy, "dy")
    B, H2, W2, Cc = dy.shape
    dx = torch.empty((B, H2 // 2, W2 // 2, Cc), dtype=torch.float32, device=dy.device)
    with _prof("nsr_nearest_up2_bwd_nhwc", (dy.numel(),), 0.0, 5.0 * dy.numel()):
        check(_lib.lib().nsr_nearest_up2_bwd_nhwc(dy.data_ptr(), dx.data_ptr(), B, H2 // 2, W2 // 2, Cc, _stream()),
              "nsr_nearest_up2_bwd_nhwc")
    _count(1)
    return dx


def mish_fwd(x: Tensor) -> Tensor:
    _chk(x, "x")
    y = torch.empty_like(x)
    with _prof("nsr_mish_fwd", (x.numel(),), 0.0, 8.0 * x.numel()):
        check(_lib.lib().nsr_mish_fwd(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "nsr_mish_fwd")
    _count(1)
    return y


def mish_bwd(dy: Tensor, x: Tensor) -> Tensor:
    _chk(dy, "dy"), _chk(x, "x")
    dx = torch.empty_like(x)
    with _prof("nsr_mish_bwd", (x.numel(),), 0.0, 12.0 * x.numel()):
        check(_lib.lib().nsr_mish_bwd(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel(), _stream()), "nsr_mish_bwd")
    _count(1)
    return dx


def mul_sigmoid_fwd(x: Tensor, s: Tensor) -> Tensor:
    _chk(x, "x"), _chk(s, "s")
    y = torch.empty_like(x)
    with _prof("nsr_mul_sigmoid_fwd", (x.numel(),), 0.0, 12.0 * x.numel()):
        check(_lib.lib().nsr_mul_sigmoid_fwd(x.data_ptr(), s.data_ptr(), y.data_ptr(), x.numel(), _stream()),
              "nsr_mul_sigmoid_fwd")
    _count(1)
    return y


def mul_sigmoid_bwd(dy: Tensor, x: Tensor, s: Tensor):
    _chk(dy, "dy"), _chk(x, "x"), _chk(s, "s")
    dx, ds = torch.empty_like(x), torch.empty_like(x)
    with _prof("nsr_mul_sigmoid_bwd", (x.numel(),), 0.0, 20.0 * x.numel()):
        check(_lib.lib().nsr_mul_sigmoid_bwd(dy.data_ptr(), x.data_ptr(), s.data_ptr(), dx.data_ptr(), ds.data_ptr(), x.numel(),
                                             _stream()), "nsr_mul_sigmoid_bwd")
    _count(1)
    return dx, ds


def add_repeat_interleave_(y: Tensor, x: Tensor, r: int) -> Tensor:
    """y[..., c*r + k] += x[..., c] in place (NHWC)."""
    _chk(y, "y"), _chk(x, "x")
    c = x.shape[-1]
    rows = x.numel() // c
    check(_lib.lib().nsr_add_repeat_interleave(y.data_ptr(), x.data_ptr(), rows, c, r, _stream()), "nsr_add_repeat_interleave")
    _count(1)
    return y


def groupnorm_fwd(x: Tensor, gamma: Tensor, beta: Tensor, groups: int, eps: float = 1e-5, residual: Tensor | None = None):
    """nn.GroupNorm on NHWC [B,H,W,C] (+ residual); returns (y, mean, rstd)."""
    for t, n in ((x, "x"), (gamma, "gamma"), (beta, "beta"), (residual, "residual")):
        _chk(t, n)
    B, H, W, Cc = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(B * groups, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    L = _lib.lib()
    ws = scratch(L.nsr_groupnorm_workspace(B, Cc, groups), x.device)
    with _prof("nsr_groupnorm_fwd", (x.numel(),), 0.0, 16.0 * x.numel()):
        check(L.nsr_groupnorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _p(residual), y.data_ptr(), mean.data_ptr(),
                                  rstd.data_ptr(), B, H * W, Cc, groups, eps, ws.data_ptr(), ws.numel(), _stream()),
              "nsr_groupnorm_fwd")
    _count(3)
    return y, mean, rstd


def groupnorm_bwd(dy: Tensor, x: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor, dgamma: Tensor, dbeta: Tensor,
                  groups: int) -> Tensor:
    for t, n in ((dy, "dy"), (x, "x"), (gamma, "gamma"), (mean, "mean"), (rstd, "rstd"), (dgamma, "dgamma"), (dbeta, "dbeta")):
        _chk(t, n)
    B, H, W, Cc = x.shape
    dx = torch.empty_like(x)
    L = _lib.lib()
    ws = scratch(L.nsr_groupnorm_workspace(B, Cc, groups), x.device)
    with _prof("nsr_groupnorm_bwd", (x.numel(),), 0.0, 24.0 * x.numel()):
        check(L.nsr_groupnorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(),
                                  dgamma.data_ptr(), dbeta.data_ptr(), B, H * W, Cc, groups, ws.data_ptr(), ws.numel(),
                                  _stream()), "nsr_groupnorm_bwd")
    _count(3)
    return dx


def bilinear_up2(x: Tensor) -> Tensor:
    """F.interpolate(scale_factor=2, mode="bilinear", align_corners=False) on NHWC."""
    _chk(x, "x")
    B, H, W, C = x.shape
    y = torch.empty((B, 2 * H, 2 * W, C), dtype=torch.float32, device=x.device)
    with _prof("nsr_bilinear_up2_nhwc", (x.numel(),), 0.0, 20.0 * x.numel()):
        check(_lib.lib().nsr_bilinear_up2_nhwc(x.data_ptr(), y.data_ptr(), B, H, W, C, _stream()), "nsr_bilinear_up2_nhwc")
    _count(1)
    return y


def bilinear_up2_bwd(dy: Tensor) -> Tensor:
    _chk(dy, "dy")
    B, H2, W2, C = dy.shape
    dx = torch.empty((B, H2 // 2, W2 // 2, C), dtype=torch.float32, device=dy.device)
    with _prof("nsr_bilinear_up2_bwd_nhwc", (dy.numel(),), 0.0, 5.0 * dy.numel()):
        check(_lib.lib().nsr_bilinear_up2_bwd_nhwc(dy.data_ptr(), dx.data_ptr(), B, H2 // 2, W2 // 2, C, _stream()),
              "nsr_bilinear_up2_bwd_nhwc")
    _count(1)
    return dx


def conv4x4s2_remap(src: Tensor, cout: int, cin: int, inverse: bool = False, out: Tensor | None = None) -> Tensor:
    """w4 [cout,cin,4,4] -> w3 [cout,4cin,3,3] (or the gradient gather back when inverse)."""
    _chk(src, "src")
    if out is None:
        out = torch.empty((cout, cin, 4, 4) if inverse else (cout, 4 * cin, 3, 3), dtype=torch.float32, device=src.device)
    _chk(out, "out")
    check(_lib.lib().nsr_conv4x4s2_remap(src.data_ptr(), out.data_ptr(), cout, cin, int(inverse), _stream()),
          "nsr_conv4x4s2_remap")
    _count(1)
    return out


def spectral_norm_fwd(w_orig: Tensor, u: Tensor, v: Tensor, w_out: Tensor, sigma: Tensor, power_iterations: int = 1,
                      eps: float = 1e-12) -> None:
    """torch.nn.utils.spectral_norm forward; u, v updated in place when power_iterations > 0."""
    for t, n in ((w_orig, "w_orig"), (u, "u"), (v, "v"), (w_out, "w_out"), (sigma, "sigma")):
        _chk(t, n)
    rows = w_orig.shape[0]
    cols = w_orig.numel() // rows
    L = _lib.lib()
    ws = scratch(L.nsr_spectral_norm_workspace(rows, cols), w_orig.device)
    check(L.nsr_spectral_norm_fwd(w_orig.data_ptr(), u.data_ptr(), v.data_ptr(), w_out.data_ptr(), sigma.data_ptr(), rows, cols,
                                  power_iterations, eps, ws.data_ptr(), ws.numel(), _stream()), "nsr_spectral_norm_fwd")
    _count(3 + 4 * power_iterations)


def spectral_norm_bwd(g_wsn: Tensor, w_sn: Tensor, u: Tensor, v: Tensor, sigma: Tensor, dw_orig: Tensor,
                      accumulate: bool = False) -> None:
    for t, n in ((g_wsn, "g_wsn"), (w_sn, "w_sn"), (u, "u"), (v, "v"), (sigma, "sigma"), (dw_orig, "dw_orig")):
        _chk(t, n)
    rows = w_sn.shape[0]
    cols = w_sn.numel() // rows
    L = _lib.lib()
    ws = scratch(L.nsr_spectral_norm_workspace(rows, cols), w_sn.device)
    check(L.nsr_spectral_norm_bwd(g_wsn.data_ptr(), w_sn.data_ptr(), u.data_ptr(), v.data_ptr(), sigma.data_ptr(),
                                  dw_orig.data_ptr(), rows, cols, int(accumulate), ws.data_ptr(), ws.numel(), _stream()),
          "nsr_spectral_norm_bwd")
    _count(2)


def prelu_bwd(dy: Tensor, pre: Tensor, slope: Tensor, dslope: Tensor) -> Tensor:
    """dx = dy * prelu'(pre); dslope (overwritten) = sum dy * min(pre, 0)."""
    _chk(dy, "dy"), _chk(pre, "pre"), _chk(slope, "slope"), _chk(dslope, "dslope")
    c = dy.shape[-1]
    rows = dy.numel() // c
    dx = torch.empty_like(dy)
    L = _lib.lib()
    ws = scratch(L.nsr_prelu_bwd_workspace(c), dy.device)
    with _prof("nsr_prelu_bwd", (rows, c), 0.0, 12.0 * dy.numel()):
        check(L.nsr_prelu_bwd(dy.data_ptr(), pre.data_ptr(), slope.data_ptr(), dx.data_ptr(), dslope.data_ptr(), rows, c,
                              ws.data_ptr(), ws.numel(), _stream()), "nsr_prelu_bwd")
    _count(2)
    return dx


def nhwc_to_nchw_add_nearest(x: Tensor, base: Tensor, scale: int) -> Tensor:
    """[B,H,W,C] NHWC + nearest-upsampled NCHW base [B,C,H/s,W/s] -> NCHW [B,C,H,W]."""
    _chk(x, "x"), _chk(base, "base")
    B, H, W, Cc = x.shape
    y = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    with _prof("nsr_nhwc_to_nchw_add_nearest", (x.numel(),), 0.0, 8.0 * x.numel()):
        check(_lib.lib().nsr_nhwc_to_nchw_add_nearest(x.data_ptr(), base.data_ptr(), y.data_ptr(), B, Cc, H, W, scale,
                                                      _stream()), "nsr_nhwc_to_nchw_add_nearest")
    _count(1)
    return y


def actgrad_mul(dy: Tensor, aux: Tensor, act: str, slope: float = 0.0, dextra: Tensor | None = None) -> Tensor:
    _chk(dy, "dy"), _chk(aux, "aux"), _chk(dextra, "dextra")
    dx = torch.empty_like(dy)
    with _prof("nsr_actgrad_mul", (dy.numel(),), 0.0, 12.0 * dy.numel()):
        check(_lib.lib().nsr_actgrad_mul(dy.data_ptr(), aux.data_ptr(), _p(dextra), dx.data_ptr(), dy.numel(), ACT[act],
                                         slope, _stream()), "nsr_actgrad_mul")
    _count(1)
    return dx


# ----------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5, sti_out: bool = False,
                  f32_out: bool = True):
    """Returns (y, mean, rstd); y is fp32, an STI, or (fp32, STI) per sti_out / f32_out."""
    _chk(x, "x"), _chk(gamma, "gamma"), _chk(beta, "beta")
    c = x.shape[-1]
    rows = x.numel() // c
    y = torch.empty_like(x) if f32_out else None
    y_sti = STI(x.shape, x.device) if sti_out else None
    if y_sti is not None:
        y_sti.ones = c % 64 != 0
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    with _prof("nsr_layernorm_fwd", (rows, c), 0.0, 8.0 * x.numel()):
        check(_lib.lib().nsr_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _p(y), mean.data_ptr(),
                                           rstd.data_ptr(), rows, c, eps, _p(y_sti), _stream()), "nsr_layernorm_fwd")
    _count(1)
    return (y if y_sti is None else (y_sti if y is None else (y, y_sti))), mean, rstd


def layernorm_bwd(dy: Tensor, x: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor, dgamma: Tensor, dbeta: Tensor,
                  dres: "Tensor | STI | None" = None, sti_out: bool = False, f32_out: bool = True,
                  deferred: "DeferredWgrads | None" = None, key=None):
    """dx = LN'(dy) + dres as fp32, as an STI, or (fp32, STI) per f32_out / sti_out; dres may be fp32 or an STI.
    deferred (+ key): dgamma / dbeta are written by `deferred.finalize()` together with the weight gradients."""
    for t, n in ((dy, "dy"), (x, "x"), (gamma, "gamma"), (mean, "mean"), (rstd, "rstd"), (dgamma, "dgamma"),
                 (dbeta, "dbeta")):
        _chk(t, n)
    res_sti = dres if isinstance(dres, STI) else None
    res_f32 = None if res_sti is not None else dres
    _chk(res_f32, "dres")
    if not (f32_out or sti_out):
        raise ValueError("layernorm_bwd: no output format selected")
    if res_sti is not None and res_sti.shape != tuple(x.shape):
        raise ValueError(f"layernorm_bwd: dres image {res_sti.shape} vs x {tuple(x.shape)}")
    c = x.shape[-1]
    rows = x.numel() // c
    dx = torch.empty_like(x) if f32_out else None
    dx_sti = STI(x.shape, x.device) if sti_out else None
    L = _lib.lib()
    need = L.nsr_layernorm_bwd_workspace(c)
    if deferred is not None:
        ws = deferred.buffers.get(key)
        if ws is None or ws.numel() < need:
            ws = deferred.buffers[key] = torch.empty(need, dtype=torch.uint8, device=x.device)
    else:
        ws = scratch(need, x.device)
    with _prof("nsr_layernorm_bwd", (rows, c), 0.0, (16.0 if dres is not None else 12.0) * x.numel()):
        check(L.nsr_layernorm_bwd2(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _p(res_f32),
                                   _p(res_sti), _p(dx), None if deferred is not None else dgamma.data_ptr(),
                                   None if deferred is not None else dbeta.data_ptr(), rows, c, ws.data_ptr(), ws.numel(),
                                   _p(dx_sti), _stream()), "nsr_layernorm_bwd")
    if deferred is not None:
        blocks = L.nsr_layernorm_bwd_blocks(rows)
        for row, out in ((0, dgamma), (1, dbeta)):
            deferred.jobs.append(dict(partial=ws.data_ptr(), dw=out.data_ptr(), dbias=None,
                                      row_map=_const_i32((row,), x.device).data_ptr(), col_map=None, splitk=blocks, p_rows=2,
                                      p_cols=c, cout=1, cin=c, bias_col=-1))
        _count(1)
    else:
        _count(2)
    return dx if dx_sti is None else (dx_sti if dx is None else (dx, dx_sti))


# ----------------------------------------------------------------------------- attention
def _attn_mma_ok(c: int, heads: int, ws: int) -> bool:
    d = c // heads
    return ws == 8 and d <= 32 and d % 2 == 0


def window_attn_fwd(qkv: Tensor, table: Tensor, heads: int, ws: int, shift: int, scale: float, sti_out: bool = False):
    """qkv [B,H,W,3C] -> [B,H,W,C] (fp32, or an STI when sti_out); shift/partition/mask/bias/softmax/PV fused."""
    _chk(qkv, "qkv"), _chk(table, "table")
    B, H, W, c3 = qkv.shape
    c = c3 // 3
    out_sti = STI((B, H, W, c), qkv.device) if sti_out else None
    need_f32 = (not sti_out) or not _attn_mma_ok(c, heads, ws)
    if out_sti is not None:
        out_sti.ones = (not need_f32) and c % 64 != 0  # the tensor-core kernel writes the image itself
    out = torch.empty((B, H, W, c), dtype=torch.float32, device=qkv.device) if need_f32 else None
    with _prof("nsr_window_attn_fwd", (B * H * W, c, heads, ws), 0.0, 4.0 * (qkv.numel() + B * H * W * c)):
        check(_lib.lib().nsr_window_attn_fwd(qkv.data_ptr(), table.data_ptr(), _p(out), B, H, W, c, heads, ws, shift,
                                             1 if shift > 0 else 0, scale, _p(out_sti), _stream()), "nsr_window_attn_fwd")
    _count(1)
    return out_sti if sti_out else out


def window_attn_bwd(qkv: Tensor, table: Tensor, dout: Tensor, dtable: Tensor, heads: int, ws: int, shift: int,
                    scale: float, sti_out: bool = False):
    _chk(qkv, "qkv"), _chk(table, "table"), _chk(dout, "dout"), _chk(dtable, "dtable")
    B, H, W, c3 = qkv.shape
    c = c3 // 3
    dqkv_sti = STI(qkv.shape, qkv.device) if sti_out else None
    need_f32 = (not sti_out) or not _attn_mma_ok(c, heads, ws)
    dqkv = torch.empty_like(qkv) if need_f32 else None
    L = _lib.lib()
    wsb = scratch(L.nsr_window_attn_bwd_workspace(heads, ws), qkv.device)
    with _prof("nsr_window_attn_bwd", (B * H * W, c, heads, ws), 0.0, 4.0 * (2 * qkv.numel() + dout.numel())):
        check(L.nsr_window_attn_bwd(qkv.data_ptr(), table.data_ptr(), dout.data_ptr(), _p(dqkv), dtable.data_ptr(),
                                    B, H, W, c, heads, ws, shift, 1 if shift > 0 else 0, scale, wsb.data_ptr(), wsb.numel(),
                                    _p(dqkv_sti), _stream()), "nsr_window_attn_bwd")
    _count(2)
    return dqkv_sti if sti_out else dqkv


def window_attn_fwd_wsti(qkv: STI, table: Tensor, c: int, heads: int, ws: int, shift: int, scale: float,
                         sti_out: bool = True, engine: str | None = None, padded_out: bool = False):
    """qkv: window-ordered, head-padded STI [B,H,W,3G] (conv_fprop(..., sti_win=(ws, shift)) with head-padded weights)
    -> attention output in natural token order, [B,H,W,c] as an STI (sti_out) or fp32.  padded_out: the STI is
    [B,H,W,G] with heads padded to 32 channels (tcgen05 kernel; `ones_col` = the channel that carries 1.0, or None)."""
    _chk(table, "table")
    B, H, W, g3 = qkv.shape
    out_sti = STI((B, H, W, g3 // 3 if padded_out else c), qkv.device) if sti_out else None
    if out_sti is not None:
        out_sti.ones = (not padded_out) and c % 64 != 0
        out_sti.ones_col = (c // heads if c // heads < 32 else None) if padded_out else None
    out = None if sti_out else torch.empty((B, H, W, c), dtype=torch.float32, device=qkv.device)
    with _prof("nsr_window_attn_wsti_fwd", (B * H * W, c, heads, ws), 0.0, 4.0 * B * H * W * (g3 + c)):
        check(_lib.lib().nsr_window_attn_wsti_fwd(qkv.data_ptr(), table.data_ptr(), _p(out), _p(out_sti), int(padded_out), B, H,
                                                  W, c, heads, ws, shift, 1 if shift > 0 else 0, scale,
                                                  ENGINE[engine or WSTI_ATTN_ENGINE], _stream()), "nsr_window_attn_wsti_fwd")
    _count(1)
    return out_sti if sti_out else out


def window_attn_bwd_wsti(qkv: STI, table: Tensor, dout: STI, dtable: Tensor, c: int, heads: int, ws: int, shift: int,
                         scale: float, sti_out: bool = True, engine: str | None = None, padded_out: bool = False,
                         deferred: "DeferredWgrads | None" = None, key=None):
    """dqkv in natural token order from the window-ordered qkv / dout images; dtable overwritten.  STI [B,H,W,3c]
    (or [B,H,W,3G] head-padded when padded_out: tcgen05 kernel) or fp32 [B,H,W,3c].
    deferred (+ key): dtable is written by `deferred.finalize()`, for all layers at once."""
    _chk(table, "table"), _chk(dtable, "dtable")
    B, H, W, g3 = qkv.shape
    dqkv_sti = STI((B, H, W, g3 if padded_out else 3 * c), qkv.device) if sti_out else None
    dqkv = None if sti_out else torch.empty((B, H, W, 3 * c), dtype=torch.float32, device=qkv.device)
    L = _lib.lib()
    need = L.nsr_window_attn_bwd_workspace(heads, ws)
    eng = ENGINE[engine or WSTI_ATTN_ENGINE]
    if deferred is not None:
        wsb = deferred.buffers.get(key)
        if wsb is None or wsb.numel() < need:
            wsb = deferred.buffers[key] = torch.empty(need, dtype=torch.uint8, device=qkv.device)
    else:
        wsb = scratch(need, qkv.device)
    with _prof("nsr_window_attn_wsti_bwd", (B * H * W, c, heads, ws), 0.0, 4.0 * B * H * W * (g3 + g3 // 3 + 3 * c)):
        check(L.nsr_window_attn_wsti_bwd(qkv.data_ptr(), table.data_ptr(), dout.data_ptr(), _p(dqkv), _p(dqkv_sti),
                                         int(padded_out and sti_out), None if deferred is not None else dtable.data_ptr(), B, H, W,
                                         c, heads, ws, shift, 1 if shift > 0 else 0, scale, eng, wsb.data_ptr(), wsb.numel(),
                                         _stream()), "nsr_window_attn_wsti_bwd")
    if deferred is not None:
        gx = L.nsr_window_attn_wsti_bwd_gx(B, H, W, c, heads, ws, int(padded_out and sti_out), eng)
        if gx <= 0 or (gx + 1) * heads * 4096 * 4 > wsb.numel():
            raise RuntimeError("window_attn_bwd_wsti: inconsistent deferred work space")
        deferred.attn_jobs.append(dict(partial=wsb.data_ptr(), dbias_table=dtable.data_ptr(), gx=gx, heads=heads, ws=ws, reserved=0))
        _count(1)
    else:
        _count(3)
    return dqkv_sti if sti_out else dqkv


# ----------------------------------------------------------------------------- losses
def _loss_ws(device) -> Tensor:
    return scratch(_lib.lib().nsr_loss_workspace(), device)


def l1_loss(pred: Tensor, target: Tensor, weight: float, loss_accum: Tensor | None, want_grad: bool = True):
    _chk(pred, "pred"), _chk(target, "target")
    dpred = torch.empty_like(pred) if want_grad else None
    val = torch.empty(1, dtype=torch.float32, device=pred.device)
    with _prof("nsr_l1_loss", (pred.numel(),), 0.0, 12.0 * pred.numel()):
        check(_lib.lib().nsr_l1_loss(pred.data_ptr(), target.data_ptr(), _p(dpred), pred.numel(), weight, _p(loss_accum),
                                     val.data_ptr(), _loss_ws(pred.device).data_ptr(), _stream()), "nsr_l1_loss")
    _count(2)
    return val, dpred


def charbonnier_loss(a: Tensor, b: Tensor, weight: float, loss_accum: Tensor | None, in_scale: float = 1.0,
                     clip_min: float = 0.0, clip_max: float = 1.0, want_grad: bool = True):
    _chk(a, "a"), _chk(b, "b")
    da = torch.empty_like(a) if want_grad else None
    val = torch.empty(1, dtype=torch.float32, device=a.device)
    with _prof("nsr_charbonnier_loss", (a.numel(),), 0.0, 12.0 * a.numel()):
        check(_lib.lib().nsr_charbonnier_loss(a.data_ptr(), b.data_ptr(), _p(da), a.numel(), in_scale, clip_min, clip_max,
                                              weight, _p(loss_accum), val.data_ptr(), _loss_ws(a.device).data_ptr(),
                                              _stream()), "nsr_charbonnier_loss")
    _count(2)
    return val, da


def bce_logits_loss(logits: Tensor, label: float, weight: float, loss_accum: Tensor | None, want_grad: bool = True):
    _chk(logits, "logits")
    dl = torch.empty_like(logits) if want_grad else None
    val = torch.empty(1, dtype=torch.float32, device=logits.device)
    with _prof("nsr_bce_logits_loss", (logits.numel(),), 0.0, 8.0 * logits.numel()):
        check(_lib.lib().nsr_bce_logits_loss(logits.data_ptr(), _p(dl), logits.numel(), label, weight, _p(loss_accum),
                                             val.data_ptr(), _loss_ws(logits.device).data_ptr(), _stream()),
              "nsr_bce_logits_loss")
    _count(2)
    return val, dl


# ----------------------------------------------------------------------------- OTF degradations
RESIZE_MODE = {"area": 0, "bilinear": 1, "bicubic": 2}


def filter2d(img: Tensor, kernel: Tensor) -> Tensor:
    """filter2D (diffjpeg.py:558-584): img [B,C,H,W], kernel [B or 1, k, k]."""
    _chk(img, "img"), _chk(kernel, "kernel")
    B, Cc, H, W = img.shape
    k = kernel.shape[-1]
    if k % 2 != 1 or kernel.shape[-2] != k:
        raise ValueError("Wrong kernel size")  # the reference's message (diffjpeg.py:571-573)
    out = torch.empty_like(img)
    with _prof("nsr_filter2d", (B, Cc, H, W, k), 2.0 * img.numel() * k * k, 8.0 * img.numel()):
        check(_lib.lib().nsr_filter2d(img.data_ptr(), kernel.data_ptr(), out.data_ptr(), B, Cc, H, W, k,
                                      kernel.shape[0], _stream()), "nsr_filter2d")
    _count(1)
    return out


def resize(img: Tensor, mode: str, *, scale_factor: float | None = None, size: tuple[int, int] | None = None) -> Tensor:
    """F.interpolate(img, scale_factor=... | size=..., mode=area|bilinear|bicubic) (align_corners False, no antialias)."""
    _chk(img, "img")
    B, Cc, H, W = img.shape
    if (scale_factor is None) == (size is None):
        raise ValueError("only one of size or scale_factor should be defined")
    if size is None:
        oh, ow = int(math.floor(float(H) * scale_factor)), int(math.floor(float(W) * scale_factor))
        rh = rw = float(_np.float32(1.0 / scale_factor))  # torch: (float)(1.0 / scale) when a scale_factor is given
    else:
        oh, ow = int(size[0]), int(size[1])
        rh, rw = float(_np.float32(H) / _np.float32(oh)), float(_np.float32(W) / _np.float32(ow))
    out = torch.empty((B, Cc, oh, ow), dtype=torch.float32, device=img.device)
    with _prof("nsr_resize", (B * Cc, H, W, oh, ow, mode), 0.0, 4.0 * (img.numel() + out.numel())):
        check(_lib.lib().nsr_resize(img.data_ptr(), out.data_ptr(), B * Cc, H, W, oh, ow, RESIZE_MODE[mode], rh, rw,
                                    _stream()), "nsr_resize")
    _count(1)
    return out


def gaussian_noise(img: Tensor, sigma: Tensor, gray: Tensor, any_gray: bool, seed: int, z: Tensor | None = None,
                   z_gray: Tensor | None = None) -> Tensor:
    """random_add_gaussian_noise_pt(clip=True) with per-sample sigma / gray flags already drawn."""
    _chk(img, "img"), _chk(sigma, "sigma"), _chk(gray, "gray"), _chk(z, "z"), _chk(z_gray, "z_gray")
    B, Cc, H, W = img.shape
    if Cc != 3:
        raise ValueError("gaussian_noise: RGB images expected")
    out = torch.empty_like(img)
    with _prof("nsr_gaussian_noise", (B, H, W), 0.0, 8.0 * img.numel()):
        check(_lib.lib().nsr_gaussian_noise(img.data_ptr(), out.data_ptr(), sigma.data_ptr(), gray.data_ptr(),
                                            int(any_gray), _p(z), _p(z_gray), B, H, W, seed & (2**64 - 1), _stream()),
              "nsr_gaussian_noise")
    _count(1)
    return out


def poisson_noise(img: Tensor, scale: Tensor, gray: Tensor, any_gray: bool, seed: int,
                  counts_color: Tensor | None = None, counts_gray: Tensor | None = None) -> Tensor:
    """random_add_poisson_noise_pt(clip=True) with per-sample scale / gray flags already drawn."""
    _chk(img, "img"), _chk(scale, "scale"), _chk(gray, "gray"), _chk(counts_color, "counts_color")
    _chk(counts_gray, "counts_gray")
    B, Cc, H, W = img.shape
    if Cc != 3:
        raise ValueError("poisson_noise: RGB images expected")
    out = torch.empty_like(img)
    wsb = _lib.lib().nsr_poisson_noise_workspace(B)
    ws = scratch(wsb, img.device)
    with _prof("nsr_poisson_noise", (B, H, W), 0.0, 12.0 * img.numel()):
        check(_lib.lib().nsr_poisson_noise(img.data_ptr(), out.data_ptr(), scale.data_ptr(), gray.data_ptr(),
                                           int(any_gray), _p(counts_color), _p(counts_gray), B, H, W,
                                           seed & (2**64 - 1), ws.data_ptr(), ws.numel(), _stream()), "nsr_poisson_noise")
    _count(2)
    return out


def jpeg(img: Tensor, quality: Tensor) -> Tensor:
    """DiffJPEG(differentiable=False)(img, quality=[B] tensor)."""
    _chk(img, "img"), _chk(quality, "quality")
    B, Cc, H, W = img.shape
    if Cc != 3 or quality.numel() != B:
        raise ValueError("jpeg: RGB images and one quality per sample expected")
    out = torch.empty_like(img)
    with _prof("nsr_jpeg", (B, H, W), 0.0, 8.0 * img.numel()):
        check(_lib.lib().nsr_jpeg(img.data_ptr(), out.data_ptr(), quality.data_ptr(), B, H, W, _stream()), "nsr_jpeg")
    _count(1)
    return out


def crop(img: Tensor, top: int, left: int, ph: int, pw: int, quantise: bool = False, out: Tensor | None = None) -> Tensor:
    _chk(img, "img")
    B, Cc, H, W = img.shape
    if out is None:
        out = torch.empty((B, Cc, ph, pw), dtype=torch.float32, device=img.device)
    with _prof("nsr_crop", (B * Cc, ph, pw), 0.0, 8.0 * out.numel()):
        check(_lib.lib().nsr_crop(img.data_ptr(), out.data_ptr(), B * Cc, H, W, top, left, ph, pw, int(quantise),
                                  _stream()), "nsr_crop")
    _count(1)
    return out


def pool_swap(pool: Tensor, new: Tensor, slots: Tensor, dequeue: bool, out: Tensor | None = None) -> Tensor | None:
    """pool[slots[i]] <-> new[i]; returns the dequeued samples (dequeue=True) or None (enqueue only)."""
    _chk(pool, "pool"), _chk(new, "new")
    if not (slots.is_cuda and slots.dtype == torch.int32 and slots.is_contiguous()):
        raise ValueError("pool_swap: slots must be a contiguous CUDA int32 tensor")
    b = new.shape[0]
    elems = new.numel() // b
    if pool.numel() // pool.shape[0] != elems:
        raise ValueError("pool_swap: sample shape mismatch")
    if dequeue and out is None:
        out = torch.empty_like(new)
    with _prof("nsr_pool_swap", (b, elems), 0.0, (12.0 if dequeue else 8.0) * new.numel()):
        check(_lib.lib().nsr_pool_swap(pool.data_ptr(), new.data_ptr(), _p(out) if dequeue else None, slots.data_ptr(),
                                       b, elems, int(dequeue), _stream()), "nsr_pool_swap")
    _count(1)
    return out if dequeue else None


# ----------------------------------------------------------------------------- MS-SSIM / consistency
def avgpool2(x: Tensor) -> Tensor:
    """F.avg_pool2d(x, 2, 2, padding=[s % 2 for s in x.shape[2:]]) (ssim_loss.py:141-143)."""
    B, Cc, H, W = x.shape
    ph, pw = H % 2, W % 2
    out = torch.empty((B, Cc, (H + 2 * ph - 2) // 2 + 1, (W + 2 * pw - 2) // 2 + 1), dtype=torch.float32, device=x.device)
    with _prof("nsr_avgpool2", (B * Cc, H, W), 0.0, 5.0 * x.numel()):
        check(_lib.lib().nsr_avgpool2(x.data_ptr(), out.data_ptr(), B * Cc, H, W, ph, pw, _stream()), "nsr_avgpool2")
    _count(1)
    return out


MSSSIM_SCALES = 5


def msssim_loss(x: Tensor, y: Tensor, window: Tensor, weight: float, c1: float, c2: float,
                loss_accum: Tensor | None = None, want_grad: bool = True):
    """mssim_loss.forward (ssim_loss.py:112-163): returns (loss [1], dloss/dx or None)."""
    _chk(x, "x"), _chk(y, "y"), _chk(window, "window")
    if x.shape != y.shape or x.dim() != 4:
        raise AssertionError(f"x: {tuple(x.shape)} and y: {tuple(y.shape)} must be the same 4-d shape")
    L = _lib.lib()
    wsz = window.shape[-1]
    dev = x.device
    xs, ys, Ps = [x], [y], []
    sums = torch.empty(2 * MSSSIM_SCALES, dtype=torch.float32, device=dev)
    counts = []
    for s in range(MSSSIM_SCALES):
        xi, yi = xs[s], ys[s]
        B, Cc, H, W = xi.shape
        counts.append(float(xi.numel()))
        P = torch.empty((3, B * Cc, H, W), dtype=torch.float32, device=dev) if want_grad else None
        ws = scratch(L.nsr_ssim_scale_workspace(B * Cc, H, W), dev)
        with _prof("nsr_ssim_scale_fwd", (B * Cc, H, W), 10.0 * wsz * wsz * xi.numel(), (8.0 + 12.0 * want_grad) * xi.numel()):
            check(L.nsr_ssim_scale_fwd(xi.data_ptr(), yi.data_ptr(), window.data_ptr(), wsz, c1, c2,
                                       int(s == MSSSIM_SCALES - 1), _p(P), sums[2 * s:].data_ptr(), B * Cc, H, W,
                                       ws.data_ptr(), ws.numel(), _stream()), "nsr_ssim_scale_fwd")
        _count(2)
        Ps.append(P)
        if s < MSSSIM_SCALES - 1:
            xs.append(avgpool2(xi))
            ys.append(avgpool2(yi))
    counts_t = _const_vec(tuple(counts), dev)
    coef = torch.empty(MSSSIM_SCALES, dtype=torch.float32, device=dev)
    val = torch.empty(1, dtype=torch.float32, device=dev)
    check(L.nsr_msssim_finalize(sums.data_ptr(), counts_t.data_ptr(), MSSSIM_SCALES, weight, coef.data_ptr(), val.data_ptr(),
                                _p(loss_accum), _stream()), "nsr_msssim_finalize")
    _count(1)
    if not want_grad:
        return val, None
    dx = None
    for s in reversed(range(MSSSIM_SCALES)):
        xi, yi = xs[s], ys[s]
        B, Cc, H, W = xi.shape
        d = torch.empty_like(xi)
        ch, cw = (dx.shape[2], dx.shape[3]) if dx is not None else (0, 0)
        with _prof("nsr_ssim_scale_bwd", (B * Cc, H, W), 6.0 * wsz * wsz * xi.numel(), 24.0 * xi.numel()):
            check(L.nsr_ssim_scale_bwd(Ps[s].data_ptr(), xi.data_ptr(), yi.data_ptr(), window.data_ptr(), wsz,
                                       coef[s:].data_ptr(), _p(dx), ch, cw, H % 2, W % 2, d.data_ptr(), B * Cc, H, W,
                                       _stream()), "nsr_ssim_scale_bwd")
        _count(1)
        dx = d
    return val, dx


_CONST_VECS: dict = {}


def _const_vec(values: tuple, device) -> Tensor:
    """Small constant fp32 device vectors, uploaded once per (values, device)."""
    key = (values, device.index)
    t = _CONST_VECS.get(key)
    if t is None:
        t = _CONST_VECS[key] = torch.tensor(values, dtype=torch.float32, device=device)
    return t


def clamp(x: Tensor, lo: float, hi: float) -> Tensor:
    _chk(x, "x")
    out = torch.empty_like(x)
    check(_lib.lib().nsr_clamp(x.data_ptr(), out.data_ptr(), x.numel(), lo, hi, _stream()), "nsr_clamp")
    _count(1)
    return out


def consistency_loss(x: Tensor, y: Tensor, blur_kernel: Tensor | None, saturation: float, brightness: float, cosim: bool,
                     weight: float, loss_accum: Tensor | None = None, want_grad: bool = True):
    """consistency_loss.forward (consistency_loss.py:146-192, criterion chc): (loss [1], dloss/dx or None)."""
    _chk(x, "x"), _chk(y, "y"), _chk(blur_kernel, "blur_kernel")
    B, Cc, H, W = x.shape
    if Cc != 3 or x.shape != y.shape:
        raise ValueError(f"Input size must have a shape of (*, 3, H, W). Got {tuple(x.shape)}")
    L, dev = _lib.lib(), x.device
    xc, yc = clamp(x, 1 / 255, 1.0), clamp(y, 1 / 255, 1.0)
    if blur_kernel is not None:
        k = blur_kernel.shape[-1]
        xb, yb = filter2d(xc, blur_kernel.view(1, k, k)), filter2d(yc, blur_kernel.view(1, k, k))
    else:
        xb, yb = xc, yc
    lx = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    ly = torch.empty_like(lx)
    val = torch.empty(1, dtype=torch.float32, device=dev)
    wsb = L.nsr_consistency_workspace(B, H, W)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)  # own buffer: fwd -> bwd state must survive other scratch users
    with _prof("nsr_consistency_fwd", (B, H, W), 0.0, 56.0 * B * H * W):
        check(L.nsr_consistency_fwd(x.data_ptr(), y.data_ptr(), xb.data_ptr(), yb.data_ptr(), saturation, brightness,
                                    int(cosim), weight, lx.data_ptr(), ly.data_ptr(), val.data_ptr(), _p(loss_accum), B, H, W,
                                    ws.data_ptr(), ws.numel(), _stream()), "nsr_consistency_fwd")
    _count(3)
    if not want_grad:
        return val, None
    g_blur, d_direct = torch.empty_like(x), torch.empty_like(x)
    with _prof("nsr_consistency_bwd", (B, H, W), 0.0, 68.0 * B * H * W):
        check(L.nsr_consistency_bwd(x.data_ptr(), y.data_ptr(), xb.data_ptr(), lx.data_ptr(), ly.data_ptr(), saturation, weight,
                                    g_blur.data_ptr(), d_direct.data_ptr(), B, H, W, ws.data_ptr(), _stream()),
              "nsr_consistency_bwd")
    _count(1)
    dx = torch.empty_like(x)
    if blur_kernel is not None:
        r = k // 2
        dpad = torch.empty((B, 3, H + 2 * r, W + 2 * r), dtype=torch.float32, device=dev)
        with _prof("nsr_corr2d_zero_ext", (B * 3, H, W, k), 2.0 * k * k * dpad.numel(), 4.0 * (x.numel() + dpad.numel())):
            check(L.nsr_corr2d_zero_ext(g_blur.data_ptr(), blur_kernel.data_ptr(), dpad.data_ptr(), B * 3, H, W, k, r, 1,
                                        _stream()), "nsr_corr2d_zero_ext")
        check(L.nsr_reflect_fold(dpad.data_ptr(), d_direct.data_ptr(), x.data_ptr(), dx.data_ptr(), B * 3, H, W, r, _stream()),
              "nsr_reflect_fold")
    else:
        check(L.nsr_reflect_fold(g_blur.data_ptr(), d_direct.data_ptr(), x.data_ptr(), dx.data_ptr(), B * 3, H, W, 0, _stream()),
              "nsr_reflect_fold")
    _count(2)
    return val, dx


# ----------------------------------------------------------------------------- HAT pieces
def xwin_attn_fwd(qkv: Tensor, table: Tensor, heads: int, ws: int, ows: int, shift: int, scale: float):
    """qkv [B,H,W,3C] -> (out [B,H,W,C], lse).  ows == ws: HAB window self-attention; ows > ws: OCAB."""
    _chk(qkv, "qkv"), _chk(table, "table")
    B, H, W, c3 = qkv.shape
    c = c3 // 3
    L = _lib.lib()
    eng = _lib.ENGINE["simt" if (DEFAULT_ENGINE == "simt" or not XWIN_TENSOR_CORES & 1) else "auto"]  # the exact-fp32 engines go together
    out = torch.empty((B, H, W, c), dtype=torch.float32, device=qkv.device)
    lse = torch.empty(L.nsr_xwin_attn_stat_floats(B, H, W, heads, ws), dtype=torch.float32, device=qkv.device)
    nk = ows * ows
    with _prof("nsr_xwin_attn_fwd", (B * H * W, c, heads, ws, ows), 4.0 * B * H * W * nk * c, 16.0 * qkv.numel() / 3):
        check(L.nsr_xwin_attn_fwd(qkv.data_ptr(), table.data_ptr(), out.data_ptr(), lse.data_ptr(), B, H, W, c, heads, ws, ows,
                                  shift, int(shift > 0), scale, eng, _stream()), "nsr_xwin_attn_fwd")
    _count(1)
    return out, lse


def xwin_attn_bwd(qkv: Tensor, table: Tensor, out: Tensor, dout: Tensor, lse: Tensor, dtable: Tensor, heads: int, ws: int,
                  ows: int, shift: int, scale: float) -> Tensor:
    _chk(qkv, "qkv"), _chk(table, "table"), _chk(out, "out"), _chk(dout, "dout"), _chk(dtable, "dtable")
    B, H, W, c3 = qkv.shape
    c = c3 // 3
    L = _lib.lib()
    eng = _lib.ENGINE["simt" if (DEFAULT_ENGINE == "simt" or not XWIN_TENSOR_CORES & 2) else "auto"]
    dqkv = torch.empty_like(qkv)
    ws_t = scratch(L.nsr_xwin_attn_bwd_workspace(B, H, W, c, heads, ws, ows), qkv.device)
    nk = ows * ows
    with _prof("nsr_xwin_attn_bwd", (B * H * W, c, heads, ws, ows), 14.0 * B * H * W * nk * c, 28.0 * qkv.numel() / 3):
        check(L.nsr_xwin_attn_bwd(qkv.data_ptr(), table.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                  dqkv.data_ptr(), dtable.data_ptr(), B, H, W, c, heads, ws, ows, shift, int(shift > 0), scale,
                                  eng, ws_t.data_ptr(), ws_t.numel(), _stream()), "nsr_xwin_attn_bwd")
    _count(4 if ows != ws else 3)
    return dqkv


def channel_mean(x: Tensor, mul: Tensor | None, scale: float) -> Tensor:
    """[B,H,W,C] -> [B,C]: scale * sum over pixels of x (* mul)."""
    _chk(x, "x"), _chk(mul, "mul")
    B, H, W, c = x.shape
    out = torch.empty((B, c), dtype=torch.float32, device=x.device)
    with _prof("nsr_channel_mean", (B, H * W, c), 0.0, (8.0 if mul is not None else 4.0) * x.numel()):
        check(_lib.lib().nsr_channel_mean(x.data_ptr(), _p(mul), out.data_ptr(), B, H * W, c, scale, _stream()), "nsr_channel_mean")
    _count(1)
    return out


def channel_gate_fwd(pooled: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor):
    B, c = pooled.shape
    cs = w1.shape[0]
    hidden = torch.empty((B, cs), dtype=torch.float32, device=pooled.device)
    gate = torch.empty((B, c), dtype=torch.float32, device=pooled.device)
    check(_lib.lib().nsr_channel_gate_fwd(pooled.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                          hidden.data_ptr(), gate.data_ptr(), B, c, cs, _stream()), "nsr_channel_gate_fwd")
    _count(1)
    return hidden, gate


def channel_scale_add_(y: Tensor, x: Tensor, gate: Tensor, alpha: float, accumulate: bool = True) -> Tensor:
    _chk(y, "y"), _chk(x, "x")
    B, H, W, c = x.shape
    with _prof("nsr_channel_scale_add", (x.numel(),), 0.0, 12.0 * x.numel()):
        check(_lib.lib().nsr_channel_scale_add(x.data_ptr(), gate.data_ptr(), y.data_ptr(), B, H * W, c, alpha, int(accumulate),
                                               _stream()), "nsr_channel_scale_add")
    _count(1)
    return y


def channel_gate_bwd(dgate, gate, hidden, pooled, w1, w2, dw1, db1, dw2, db2) -> Tensor:
    B, c = gate.shape
    dpooled = torch.empty_like(gate)
    check(_lib.lib().nsr_channel_gate_bwd(dgate.data_ptr(), gate.data_ptr(), hidden.data_ptr(), pooled.data_ptr(), w1.data_ptr(),
                                          w2.data_ptr(), dpooled.data_ptr(), dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(),
                                          db2.data_ptr(), B, c, w1.shape[0], _stream()), "nsr_channel_gate_bwd")
    _count(1)
    return dpooled


def channel_scale_bwd(g: Tensor, gate: Tensor, dpooled: Tensor, alpha: float) -> Tensor:
    _chk(g, "g")
    B, H, W, c = g.shape
    dx = torch.empty_like(g)
    with _prof("nsr_channel_scale_bwd", (g.numel(),), 0.0, 8.0 * g.numel()):
        check(_lib.lib().nsr_channel_scale_bwd(g.data_ptr(), gate.data_ptr(), dpooled.data_ptr(), dx.data_ptr(), B, H * W, c, alpha,
                                               _stream()), "nsr_channel_scale_bwd")
    _count(1)
    return dx


# ----------------------------------------------------------------------------- augmentations
def resize_aa(src: Tensor, mode: str, *, scale_factor: float | None = None, size=None, perm: Tensor | None = None,
              dst: Tensor | None = None, top: int = 0, left: int = 0) -> Tensor:
    """clamp(F.interpolate(src[perm], ..., mode=bilinear|bicubic, antialias=True), 0, 1), optionally pasted into the
    window (top, left) of `dst`."""
    _chk(src, "src"), _chk(dst, "dst")
    B, Cc, H, W = src.shape
    if size is None:
        oh, ow = int(math.floor(float(H) * scale_factor)), int(math.floor(float(W) * scale_factor))
        rh = rw = float(_np.float32(1.0 / scale_factor))
    else:
        oh, ow = int(size[0]), int(size[1])
        rh, rw = float(_np.float32(H) / _np.float32(oh)), float(_np.float32(W) / _np.float32(ow))
    if dst is None:
        dst = torch.empty((B, Cc, oh, ow), dtype=torch.float32, device=src.device)
    with _prof("nsr_resize_aa", (B * Cc, H, W, oh, ow), 0.0, 4.0 * (src.numel() + B * Cc * oh * ow)):
        check(_lib.lib().nsr_resize_aa(src.data_ptr(), dst.data_ptr(), _p(perm), B, Cc, H, W, oh, ow, dst.shape[2], dst.shape[3],
                                       top, left, int(mode == "bicubic"), rh, rw, _stream()), "nsr_resize_aa")
    _count(1)
    return dst


def batch_mix(a: Tensor, other: Tensor, perm: Tensor | None, mode: str, lam: float = 0.0, box=(0, 0, 0, 0)) -> Tensor:
    """mode 'mixup': lam*a + (1-lam)*other[perm]; mode 'box': a with [:, :, y0:y1, x0:x1] taken from other[perm]."""
    _chk(a, "a"), _chk(other, "other")
    B, Cc, H, W = a.shape
    dst = torch.empty_like(a)
    with _prof("nsr_batch_mix", (a.numel(),), 0.0, 12.0 * a.numel()):
        check(_lib.lib().nsr_batch_mix(a.data_ptr(), other.data_ptr(), dst.data_ptr(), _p(perm), B, Cc, H, W,
                                       0 if mode == "mixup" else 1, lam, float(_np.float32(1.0 - lam)), *[int(v) for v in box],
                                       _stream()), "nsr_batch_mix")
    _count(1)
    return dst
