"""Learning-rate schedules of the three reference recipes."""
from __future__ import annotations

import bisect
from typing import Sequence

import numpy as np


class WarmupStepLR:
    """Linear warm-up from ``base_lr`` to ``peak_lr`` over ``warmup_iters`` then x``gamma`` at each
    milestone (iterations).  ResNet-18 (mix.py:181-198): 0.1 -> 1.6 over 5 epochs, x0.1 after
    epochs 40 and 80.  ResNet-50 (main.py:237-252): 0.1 -> 3.2 over 5 epochs, x0.1 at 30/60/80."""

    def __init__(self, base_lr: float, peak_lr: float, warmup_iters: int, milestones: Sequence[int], gamma: float = 0.1):
        self.base_lr, self.peak_lr, self.warmup_iters = base_lr, peak_lr, max(int(warmup_iters), 0)
        self.milestones, self.gamma = sorted(int(m) for m in milestones), gamma

    def __call__(self, it: int) -> float:
        if it < self.warmup_iters:
            return self.base_lr + (self.peak_lr - self.base_lr) * it / max(self.warmup_iters, 1)
        return self.peak_lr * self.gamma ** bisect.bisect_right(self.milestones, it)


class PiecewiseLinear:
    """DavidNet schedule (dawn.py:65; utils.py:299-301): linear interpolation through knots (in epochs)."""

    def __init__(self, knots: Sequence[float], vals: Sequence[float]):
        self.knots, self.vals = list(knots), list(vals)

    def __call__(self, t: float) -> float:
        return float(np.interp([t], self.knots, self.vals)[0])


class IterLRScheduler:
    """Step schedule indexed by iteration with explicit multipliers (train_util.py:68-107):
    lr is multiplied by ``lr_mults[i]`` when iteration ``milestones[i]`` is reached."""

    def __init__(self, base_lr: float, milestones: Sequence[int], lr_mults: Sequence[float], last_iter: int = -1):
        assert len(milestones) == len(lr_mults), "milestones and lr_mults must have the same length"
        self.base_lr, self.milestones, self.lr_mults = base_lr, list(milestones), list(lr_mults)
        self.last_iter = last_iter

    def get_lr(self, it: int) -> float:
        lr = self.base_lr
        for m, k in zip(self.milestones, self.lr_mults):
            if it >= m:
                lr *= k
        return lr

    def step(self, this_iter=None) -> float:
        self.last_iter = self.last_iter + 1 if this_iter is None else this_iter
        return self.get_lr(self.last_iter)

    __call__ = get_lr


def linear_warmup(lr: float, step: int, warmup_iters: int) -> float:
    """Extra warm-up of the DavidNet trainer (utils.py:317-325)."""
    return lr if step > warmup_iters else lr * step / max(warmup_iters, 1)
