"""Device-resident multi-tensor table (NsrParamEntry[]) shared by the fused optimizers."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from .._lib import OPT_CHUNK, NsrParamEntry


class ParamTable:
    def __init__(self):
        self.key = None
        self.dev = None
        self.n = 0
        self.chunks = 0

    def build(self, rows: list) -> "ParamTable":
        """rows: list of dicts with tensors p,g,exp_avg,exp_avg_sq,exp_avg_diff,z,neg_pre_grad,ema (None ok)."""
        fields = ("p", "g", "exp_avg", "exp_avg_sq", "exp_avg_diff", "z", "neg_pre_grad", "ema")
        key = tuple(tuple(0 if r.get(f) is None else r[f].data_ptr() for f in fields) + (r["p"].numel(),) for r in rows)
        if key == self.key:
            return self
        arr = (NsrParamEntry * len(rows))()
        base = 0
        for i, r in enumerate(rows):
            for f in fields:
                t = r.get(f)
                if t is not None and not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                    raise ValueError(f"fused optimizer: tensor '{f}' must be contiguous CUDA fp32")
                setattr(arr[i], f, None if t is None else t.data_ptr())
            n = r["p"].numel()
            arr[i].n = n
            arr[i].chunk_base = base
            base += (n + OPT_CHUNK - 1) // OPT_CHUNK
        raw = bytes(arr)
        host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
        self.dev = host.to(rows[0]["p"].device)
        self.key, self.n, self.chunks = key, len(rows), base
        return self


def grad_sumsq(table: ParamTable, out: torch.Tensor) -> None:
    L = _lib.lib()
    from ..ops import _stream, scratch  # noqa: PLC0415
    ws = scratch(L.nsr_grad_sumsq_workspace(), out.device)
    _lib.check(L.nsr_grad_sumsq(table.dev.data_ptr(), table.n, table.chunks, out.data_ptr(), ws.data_ptr(), _stream()),
               "nsr_grad_sumsq")
    from .. import ops as ops_mod  # noqa: PLC0415
    ops_mod._count(2)
