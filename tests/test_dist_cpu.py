"""Host-side data-parallel logic on CPU with the gloo backend (world_size 2):
  * EnlargedSampler shards are disjoint, equal-sized and rank-strided;
  * mean all-reduce of per-rank gradients == gradient of the single-process big batch (the
    multi-GPU parity definition), checked on the oracle."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neosr_b200.data import EnlargedSampler, SyntheticPairedDataset


def test_sampler_shards_are_disjoint_and_strided():
    ds = SyntheticPairedDataset(length=37)
    shards = []
    for r in range(4):
        s = EnlargedSampler(ds, num_replicas=4, rank=r, ratio=3)
        s.set_epoch(5)
        shards.append(list(iter(s)))
    assert len({len(s) for s in shards}) == 1
    g = torch.Generator()
    g.manual_seed(5)
    perm = [v % 37 for v in torch.randperm(len(shards[0]) * 4, generator=g).tolist()]
    for r in range(4):
        assert shards[r] == perm[r::4]
    item = ds[3]
    assert item["lq"].shape == (3, 64, 64) and item["gt"].shape == (3, 256, 256)
    assert torch.equal(item["gt"], torch.round(item["gt"] * 255) / 255)


def _worker(rank: int, world: int, initfile: str, out: dict):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    try:
        from neosr_b200.dist import allreduce_mean_, get_dist_info
        from oracle.swinir import SwinIRConfig, swinir_forward, swinir_param_shapes, synth_params
        assert get_dist_info() == (rank, world)
        cfg = SwinIRConfig(img_size=16, embed_dim=24, depths=(2,), num_heads=(2,), window_size=8, upscale=2)
        p = {k: v.requires_grad_(True) for k, v in synth_params(swinir_param_shapes(cfg), seed=3).items()}
        g = torch.Generator().manual_seed(11)
        lq = torch.rand(4, 3, 16, 16, generator=g)
        gt = torch.rand(4, 3, 32, 32, generator=g)
        sl = slice(rank * 2, rank * 2 + 2)  # per-rank batch = 2 of the global 4
        loss = ((swinir_forward(p, cfg, lq[sl]) - gt[sl]) ** 2).mean()
        grads = torch.autograd.grad(loss, list(p.values()))
        flat = torch.cat([x.reshape(-1) for x in grads])
        allreduce_mean_(flat)
        if rank == 0:
            big = ((swinir_forward(p, cfg, lq) - gt) ** 2).mean()
            ref = torch.cat([x.reshape(-1) for x in torch.autograd.grad(big, list(p.values()))])
            out["err"] = float((flat - ref).abs().max() / ref.abs().max())
    finally:
        dist.destroy_process_group()


def test_two_rank_mean_allreduce_equals_big_batch():
    with tempfile.TemporaryDirectory() as d:
        mgr = mp.Manager()
        out = mgr.dict()
        mp.spawn(_worker, args=(2, os.path.join(d, "init"), out), nprocs=2, join=True)
        assert out["err"] < 1e-5, out["err"]
