"""Data-parallel plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch on
the GPU box, gloo in CPU tests).  The hot path has exactly one collective per network per step:
an all-reduce (mean) of the flat fp32 gradient buffer (what DDP's bucket reducer
does in the reference, base.py:140-146).  There is no activation or parameter sharding."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_dist(launcher: str = "pytorch", backend: str = "nccl") -> tuple[int, int]:
    """neosr/utils/dist_util.py:12-29 (`pytorch` launcher: RANK / WORLD_SIZE / LOCAL_RANK from the env)."""
    if launcher != "pytorch":
        raise NotImplementedError("only the torchrun ('pytorch') launcher is supported")
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend=backend)
    return rank, dist.get_world_size()


def get_dist_info() -> tuple[int, int]:
    """neosr/utils/dist_util.py:65-72."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    """In-place mean over ranks of one flat buffer (gradients, or the loss-log vector)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if flat.is_cuda:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:  # gloo has no AVG
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(dist.get_world_size())
    return flat


def broadcast_params_(params, src: int = 0) -> None:
    """Make replicas identical at start-up (DDP broadcasts rank-0 parameters when wrapping)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for p in params:
            dist.broadcast(p.data, src)
