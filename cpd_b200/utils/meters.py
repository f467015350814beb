"""Meters, accuracy, device-side timing (reference: train_util.py:21-65, DavidNet/utils.py:28-56;
the reference only has host wall-clock timers — here throughput is CUDA-event timed)."""
from __future__ import annotations

import time
from collections import deque
from typing import List

import torch
import torch.distributed as dist


class AverageMeter:
    """Windowed (``length`` > 0) or global running average."""

    def __init__(self, length: int = 0):
        self.length = length
        self.reset()

    def reset(self):
        self.history = deque(maxlen=self.length) if self.length > 0 else None
        self.count, self.sum, self.val, self.avg = 0, 0.0, 0.0, 0.0

    def update(self, val, n: int = 1):
        self.val = float(val)
        if self.history is not None:
            self.history.append(self.val)
            self.avg = sum(self.history) / len(self.history)
        else:
            self.sum += self.val * n
            self.count += n
            self.avg = self.sum / max(self.count, 1)


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)) -> List[torch.Tensor]:
    """Top-k precision in percent."""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1))
    return [correct[:k].reshape(-1).float().sum() * (100.0 / target.size(0)) for k in topk]


class Timer:
    """Split timer: ``t()`` returns seconds since the previous call; synchronises the device first."""

    def __init__(self, sync: bool = True):
        self.sync = sync and torch.cuda.is_available()
        self.times = [self._now()]
        self.total_time = 0.0

    def _now(self):
        if self.sync:
            torch.cuda.synchronize()
        return time.perf_counter()

    def __call__(self, include_in_total: bool = True) -> float:
        self.times.append(self._now())
        dt = self.times[-1] - self.times[-2]
        if include_in_total:
            self.total_time += dt
        return dt


class DeviceTimer:
    """CUDA-event timing of a region, reported as the MAX over ranks (the BASELINE metric)."""

    def __init__(self, device=None):
        self.cuda = torch.cuda.is_available()
        self.device = device
        if self.cuda:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
        self._t0 = 0.0

    def start(self):
        if self.cuda:
            torch.cuda.synchronize()
            self.e0.record()
        else:
            self._t0 = time.perf_counter()

    def stop_ms(self, reduce_max: bool = True) -> float:
        if self.cuda:
            self.e1.record()
            torch.cuda.synchronize()
            ms = self.e0.elapsed_time(self.e1)
        else:
            ms = (time.perf_counter() - self._t0) * 1e3
        if reduce_max and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda" if self.cuda else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms


class Metric:
    """Running mean of a scalar averaged over ranks (main.py:273-285)."""

    def __init__(self, name: str):
        self.name, self.sum, self.n = name, 0.0, 0

    def update(self, val):
        v = val.detach().float().clone() if isinstance(val, torch.Tensor) else torch.tensor(float(val))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            if v.device.type == "cpu" and dist.get_backend() == "nccl":
                v = v.cuda()
            dist.all_reduce(v)
            v = v / dist.get_world_size()
        self.sum += float(v)
        self.n += 1

    @property
    def avg(self) -> float:
        return self.sum / max(self.n, 1)
