"""Device-side batch prefetcher with the interface of the reference's `CUDAPrefetcher`
(neosr/data/prefetch_dataloader.py:69-113: `next()` hands out a batch whose tensors are already on the GPU and starts the
copy of the following one; `reset()` restarts the epoch), which train.py selects with `prefetch_mode = "cuda"`.

B200-side differences: host tensors that are not pinned are staged through a reusable pinned buffer (an unpinned source makes
`non_blocking=True` a synchronous copy), and the copy stream is created once per prefetcher.

Measured on the C3 step (26.7 MB per batch, loss read back every iteration): no gain over copying inside `feed_data` -
91.2 vs 91.7 ms per step - because that copy costs ~1 ms at the head of a 90 ms step while `next()` costs the host 1 - 3 ms that
the per-step loss read exposes; `bench.py`'s end-to-end leg therefore feeds pinned host batches straight to `feed_data`.  The
class exists for the reference's `prefetch_mode = "cuda"` train loop (train.py:255-262)."""
from __future__ import annotations

from typing import Any, Iterable

import torch


class CUDAPrefetcher:
    def __init__(self, loader: Iterable[dict], opt: dict[str, Any] | None = None, device: torch.device | str = "cuda"):
        self.ori_loader = loader
        self.loader = iter(loader)
        self.opt = opt or {}
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("neosr_b200.CUDAPrefetcher stages batches on a CUDA device; there is no CPU path")
        self.stream = torch.cuda.Stream(device=self.device)
        self._pinned: dict = {}
        self.batch: dict | None = None
        self.preload()

    def _pin(self, key: str, t: torch.Tensor) -> torch.Tensor:
        if t.is_pinned() or t.is_cuda:
            return t
        buf = self._pinned.get(key)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._pinned[key] = buf
        # the previous copy out of this buffer was enqueued on self.stream: wait for it before overwriting the source
        self.stream.synchronize()
        buf.copy_(t)
        return buf

    def preload(self) -> None:
        try:
            batch = next(self.loader)
        except StopIteration:
            self.batch = None
            return
        out = dict(batch)
        with torch.cuda.stream(self.stream):
            for k, v in batch.items():
                if torch.is_tensor(v):
                    out[k] = self._pin(k, v).to(device=self.device, non_blocking=True)
        self.batch = out

    def next(self) -> dict | None:
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.stream)
        batch = self.batch
        if batch is not None:
            for v in batch.values():
                if torch.is_tensor(v) and v.is_cuda:
                    v.record_stream(cur)  # allocated on the copy stream, consumed on the compute stream
        self.preload()
        return batch

    def reset(self) -> None:
        self.loader = iter(self.ori_loader)
        self.preload()
