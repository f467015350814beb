"""Synthetic datasets of the named shapes (there is no network / dataset on the box):
CIFAR-shape 3x32x32 / 10 classes, ImageNet-shape 3x224x224 / 1000 classes.  Images are
generated once in pinned host memory; every step's batch is a real host->device copy."""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import torch


class SyntheticImageDataset(torch.utils.data.Dataset):
    """Random images / labels of a named shape (no dataset download in the sandbox); ``learnable=True`` makes the
    label a function of the image so that accuracy can rise."""
    def __init__(self, n: int, shape=(3, 32, 32), num_classes: int = 10, seed: int = 0, learnable: bool = True):
        g = torch.Generator().manual_seed(seed)
        self.labels = torch.randint(0, num_classes, (n,), generator=g)
        self.images = torch.randn(n, *shape, generator=g)
        if learnable:      # plant a class-dependent mean so that training can actually reduce the loss
            proto = torch.randn(num_classes, *shape, generator=g) * 0.5
            self.images += proto[self.labels]
        self.num_classes = num_classes

    def __len__(self):
        return self.labels.numel()

    def __getitem__(self, i):
        return self.images[i], self.labels[i]


def cifar_shape(n: int = 50000, seed: int = 0) -> SyntheticImageDataset:
    return SyntheticImageDataset(n, (3, 32, 32), 10, seed)


def imagenet_shape(n: int = 1024, seed: int = 0) -> SyntheticImageDataset:
    return SyntheticImageDataset(n, (3, 224, 224), 1000, seed)


class PinnedBatchStream:
    """Endless stream of (images, labels) batches in PINNED host memory, cycling over a small
    pool — the host side of the end-to-end benchmark (one real H2D copy per step)."""

    def __init__(self, batch: int, shape=(3, 32, 32), num_classes: int = 10, pool: int = 4, seed: int = 0,
                 pin: Optional[bool] = None):
        g = torch.Generator().manual_seed(seed)
        pin = torch.cuda.is_available() if pin is None else pin
        self.batches = []
        for _ in range(pool):
            x = torch.randn(batch, *shape, generator=g)
            y = torch.randint(0, num_classes, (batch,), generator=g)
            if pin:
                x, y = x.pin_memory(), y.pin_memory()
            self.batches.append((x, y))
        self.i = 0
        self.bytes_per_batch = self.batches[0][0].numel() * 4 + self.batches[0][1].numel() * 8

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        return self

    def __next__(self):
        b = self.batches[self.i % len(self.batches)]
        self.i += 1
        return b


class Batches:
    """DataLoader wrapper of the DavidNet recipe (reference: example/DavidNet/utils.py:152-178): yields
    {'input', 'target'} dicts already on the device; uses a DistributedSampler when ``distributed``."""

    def __init__(self, dataset, batch_size, shuffle, num_workers=0, drop_last=False, distributed=False, device=None):
        self.dataset = dataset
        self.device = device or (torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu"))
        sampler = None
        if distributed:
            from torch.utils.data.distributed import DistributedSampler as TorchDS
            sampler = TorchDS(dataset, shuffle=shuffle)
        self.loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, num_workers=num_workers, pin_memory=torch.cuda.is_available(),
                                                  shuffle=(shuffle and sampler is None), drop_last=drop_last, sampler=sampler)

    def __iter__(self):
        return ({"input": x.to(self.device, non_blocking=True), "target": y.to(self.device, non_blocking=True).long()} for (x, y) in self.loader)

    def __len__(self):
        return len(self.loader)
