"""Optimizer-step hook for runner-style frameworks.

The reference's FCN/Cityscapes workload lives in an out-of-tree mmcv fork where APS is switched on by
editing ``mmcv/runner/hooks/optimizer.py`` (README.md:132-151).  This is that integration point as a
library feature: an object with mmcv's ``OptimizerHook`` interface (``after_train_iter(runner)``) that
inserts the custom-precision gradient all-reduce between backward and ``optimizer.step()``; it also works
standalone through ``step(model, optimizer, loss)``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from ..config import PrecisionConfig
from .api import sum_gradients


class APSOptimizerHook:
    """Optimizer hook in the style of the reference's mmcv fork (README.md:133-150): backward, low-precision
    ``sum_gradients`` (APS / Kahan), optional gradient clipping, ``optimizer.step()``.  ``after_train_iter(runner)`` makes
    it usable as an mmcv runner hook; ``step(model, optimizer, loss)`` works with any training loop."""
    def __init__(self, grad_exp: int = 5, grad_man: int = 2, use_APS: bool = True, use_kahan: bool = False,
                 grad_clip: Optional[dict] = None, precision: Optional[PrecisionConfig] = None):
        self.cfg = precision or PrecisionConfig(grad_exp=grad_exp, grad_man=grad_man, use_aps=use_APS, use_kahan=use_kahan)
        self.grad_clip = grad_clip

    def clip_grads(self, params):
        params = [p for p in params if p.requires_grad and p.grad is not None]
        if params and self.grad_clip:
            return torch.nn.utils.clip_grad_norm_(params, **self.grad_clip)
        return None

    def step(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss: torch.Tensor):
        optimizer.zero_grad()
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        (loss / world).backward()                     # the low-precision SUM over ranks is then the mean gradient
        sum_gradients(model, use_APS=self.cfg.use_aps, grad_exp=self.cfg.grad_exp, grad_man=self.cfg.grad_man,
                      use_kahan=self.cfg.use_kahan)
        norm = self.clip_grads(model.parameters())
        optimizer.step()
        return norm

    # mmcv.runner.Hook interface --------------------------------------------------------------------
    def after_train_iter(self, runner):
        norm = self.step(runner.model, runner.optimizer, runner.outputs["loss"])
        if norm is not None and hasattr(runner, "log_buffer"):
            runner.log_buffer.update({"grad_norm": float(norm)}, runner.outputs.get("num_samples", 1))
