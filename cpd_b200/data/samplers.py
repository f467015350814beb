"""Samplers (reference: example/ResNet18/utils/train_util.py:110-265)."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist
from torch.utils.data.sampler import Sampler


def _world(world_size, rank):
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    return world_size, rank


class GivenIterationSampler(Sampler):
    """Exactly ``total_iter * batch_size`` indices from a seed-0 shuffle of the tiled dataset;
    resumable from ``last_iter``; single use."""

    def __init__(self, dataset, total_iter: int, batch_size: int, last_iter: int = -1, seed: int = 0):
        self.dataset, self.total_iter, self.batch_size, self.last_iter, self.seed = dataset, total_iter, batch_size, last_iter, seed
        self.total_size = total_iter * batch_size
        self.indices = self._gen()
        self.used = False

    def _gen(self):
        n = len(self.dataset)
        rng = np.random.RandomState(self.seed)
        idx = np.tile(np.arange(n), math.ceil(self.total_size / n))[: self.total_size]
        rng.shuffle(idx)
        return idx

    def __iter__(self):
        if self.used:
            raise RuntimeError("this sampler is not designed to be called more than once")
        self.used = True
        return iter(self.indices[(self.last_iter + 1) * self.batch_size:])

    def __len__(self):
        return self.total_size - (self.last_iter + 1) * self.batch_size


class DistributedGivenIterationSampler(Sampler):
    """Rank slice of a seed-0 shuffle of ``total_iter * batch_size * world`` tiled indices
    (train_util.py:159-222).  Unlike the reference, explicit ``world_size=`` / ``rank=`` are honoured."""

    def __init__(self, dataset, total_iter: int, batch_size: int, world_size=None, rank=None, last_iter: int = -1, seed: int = 0):
        self.world_size, self.rank = _world(world_size, rank)
        assert self.rank < self.world_size
        self.dataset, self.total_iter, self.batch_size, self.last_iter, self.seed = dataset, total_iter, batch_size, last_iter, seed
        self.total_size = total_iter * batch_size
        self.indices = self._gen()
        self.used = False

    def _gen(self):
        n = len(self.dataset)
        all_size = self.total_size * self.world_size
        rng = np.random.RandomState(self.seed)
        idx = np.tile(np.arange(n), math.ceil(all_size / n))[:all_size]
        rng.shuffle(idx)
        return idx[self.rank * self.total_size:(self.rank + 1) * self.total_size]

    def __iter__(self):
        if self.used:
            raise RuntimeError("this sampler is not designed to be called more than once")
        self.used = True
        return iter(self.indices[(self.last_iter + 1) * self.batch_size:])

    def __len__(self):
        return self.total_size - (self.last_iter + 1) * self.batch_size


class DistributedSampler(Sampler):
    """Per-epoch deterministic shuffle, rank-strided split, optional ``round_up`` padding
    (train_util.py:225-265)."""

    def __init__(self, dataset, world_size=None, rank=None, round_up: bool = True, shuffle: bool = True):
        self.world_size, self.rank = _world(world_size, rank)
        self.dataset, self.round_up, self.shuffle, self.epoch = dataset, round_up, shuffle, 0
        self.num_samples = int(math.ceil(len(dataset) / self.world_size))
        self.total_size = self.num_samples * self.world_size if round_up else len(dataset)

    def __iter__(self):
        if self.shuffle:
            g = torch.Generator().manual_seed(self.epoch)
            idx = torch.randperm(len(self.dataset), generator=g).tolist()
        else:
            idx = list(range(len(self.dataset)))
        if self.round_up:
            idx += idx[: self.total_size - len(idx)]
        idx = idx[self.rank:self.total_size:self.world_size]
        if self.round_up or len(idx) == self.num_samples:
            pass
        return iter(idx)

    def __len__(self):
        return self.num_samples

    def set_epoch(self, epoch: int):
        self.epoch = epoch
