"""numpy preprocessing and DAWNBench-style augmentation with per-epoch pre-drawn choices
(reference: example/DavidNet/utils.py:69-145)."""
from __future__ import annotations

import numpy as np

CIFAR_MEAN = (0.4914, 0.4822, 0.4465)
CIFAR_STD = (0.2471, 0.2435, 0.2616)


def normalise(x, mean=CIFAR_MEAN, std=CIFAR_STD):
    x = np.asarray(x, dtype=np.float32)
    return (x - np.array(mean, np.float32) * 255) / (np.array(std, np.float32) * 255)


def pad(x, border: int = 4):
    return np.pad(x, [(0, 0), (border, border), (border, border), (0, 0)], mode="reflect")


def transpose(x, source="NHWC", target="NCHW"):
    return x.transpose([source.index(d) for d in target])


class Crop:
    """Random h x w crop with pre-drawn per-sample offsets (utils.py:89-101)."""
    def __init__(self, h, w):
        self.h, self.w = h, w

    def __call__(self, x, x0, y0):
        return x[:, y0:y0 + self.h, x0:x0 + self.w]

    def options(self, x_shape):
        _c, H, W = x_shape
        return {"x0": range(W + 1 - self.w), "y0": range(H + 1 - self.h)}

    def output_shape(self, x_shape):
        return (x_shape[0], self.h, self.w)


class FlipLR:
    """Random left-right flip with a pre-drawn per-sample choice (utils.py:104-112)."""
    def __call__(self, x, choice):
        return x[:, :, ::-1].copy() if choice else x

    def options(self, x_shape):
        return {"choice": [True, False]}


class Cutout:
    """Zeroes a random h x w patch, offsets pre-drawn per sample (utils.py:115-128)."""
    def __init__(self, h, w):
        self.h, self.w = h, w

    def __call__(self, x, x0, y0):
        x = x.copy()
        x[:, y0:y0 + self.h, x0:x0 + self.w] = 0.0
        return x

    def options(self, x_shape):
        _c, H, W = x_shape
        return {"x0": range(W + 1 - self.w), "y0": range(H + 1 - self.h)}


class Transform:
    """Dataset wrapper applying a list of augmentations whose random choices are drawn once per
    epoch with ``set_random_choices`` (so workers stay deterministic)."""

    def __init__(self, dataset, transforms, seed: int = 0):
        self.dataset, self.transforms = dataset, transforms
        self.choices = None
        self.rng = np.random.RandomState(seed)

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, i):
        data, label = self.dataset[i]
        data = np.asarray(data)
        for ch, f in zip(self.choices, self.transforms):
            data = f(data, **{k: v[i] for k, v in ch.items()})
        return data, label

    def set_random_choices(self):
        self.choices = []
        shape = np.asarray(self.dataset[0][0]).shape
        n = len(self)
        for t in self.transforms:
            opts = t.options(shape)
            shape = t.output_shape(shape) if hasattr(t, "output_shape") else shape
            self.choices.append({k: self.rng.choice(np.asarray(list(v)), size=n) for k, v in opts.items()})
