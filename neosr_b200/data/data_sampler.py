"""Rank-strided sampler with the reference's sharding contract (neosr/data/data_sampler.py:8-54):
every rank draws the same epoch-seeded permutation of `ceil(len*ratio/world)*world` indices and
keeps `perm[rank::world]`.  The reference draws the permutation with a CUDA generator
(data_sampler.py:37-39); here the generator lives on the CPU so the shard assignment is
reproducible without a GPU (the sharding rule — disjoint, equal-sized, rank-strided — is what
the training step depends on, not the particular permutation)."""
from __future__ import annotations

import math
from collections.abc import Iterator

import torch
from torch.utils.data.sampler import Sampler


class EnlargedSampler(Sampler):
    def __init__(self, dataset, num_replicas: int = 1, rank: int = 0, ratio: int = 1) -> None:
        self.dataset = dataset
        self.num_replicas, self.rank, self.epoch = num_replicas, rank, 0
        self.num_samples = math.ceil(len(self.dataset) * ratio / self.num_replicas)
        self.total_size = self.num_samples * self.num_replicas

    def __iter__(self) -> Iterator[int]:
        g = torch.Generator()
        g.manual_seed(self.epoch)
        indices = torch.randperm(self.total_size, generator=g).tolist()
        n = len(self.dataset)
        indices = [v % n for v in indices]
        indices = indices[self.rank:self.total_size:self.num_replicas]
        assert len(indices) == self.num_samples
        return iter(indices)

    def __len__(self) -> int:
        return self.num_samples

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch
