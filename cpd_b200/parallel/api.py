"""Reference-compatible distributed API (CPDtorch/utils/dist_util.py) on the new engine.

``sum_gradients`` / ``normal_sum_gradients`` / ``kahan_sum_gradients`` /
``broadcast_params`` / ``DistModule`` keep the reference signatures.  Under the hood a
model's gradients are adopted into a flat arena once and reduced by ``GradReducer``
(one fused kernel per bucket) instead of the per-parameter python loops.
"""
from __future__ import annotations

import weakref
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist
from torch.nn import Module

from ..config import PrecisionConfig
from .arena import ModelArena
from .reducer import GradReducer

_CACHE: "weakref.WeakKeyDictionary[Module, Dict[Tuple, Tuple[ModelArena, GradReducer]]]" = weakref.WeakKeyDictionary()


def _engine(model: Module, cfg: PrecisionConfig) -> Tuple[ModelArena, GradReducer]:
    root = model.module if isinstance(model, DistModule) else model
    per_model = _CACHE.setdefault(root, {})
    bound = getattr(root, "_cpd_arena", None)
    key = (cfg.grad_exp, cfg.grad_man, cfg.use_aps, cfg.use_kahan, cfg.algo, cfg.wire, cfg.rounding, id(bound))
    if key not in per_model:
        arena = bound if bound is not None else _grad_arena(root, per_model)
        per_model[key] = (arena, GradReducer(arena, cfg))
    return per_model[key]


def _grad_arena(root: Module, per_model) -> ModelArena:
    for (arena, _r) in per_model.values():
        return arena                      # one gradient arena per model, shared by all configs
    return ModelArena(list(root.parameters()), rehome_params=False, bind_grads=True)


def sum_gradients(model: Module, use_APS: bool = False, grad_exp: int = 5, grad_man: int = 2,
                  use_kahan: bool = False, **extra) -> None:
    """Sum ``param.grad`` over all ranks with every intermediate cast to (grad_exp, grad_man);
    with ``use_APS`` each tensor is first shifted by a power of two chosen from the global max
    exponent (dist_util.py:22-51).  Works at world_size 1 and on CPU (Gloo) too."""
    cfg = PrecisionConfig(grad_exp=grad_exp, grad_man=grad_man, use_aps=use_APS, use_kahan=use_kahan, **extra)
    arena, reducer = _engine(model, cfg)
    arena.adopt_grads()
    reducer.reduce()


def normal_sum_gradients(model: Module, grad_exp: int = 8, grad_man: int = 23) -> None:
    sum_gradients(model, use_APS=False, grad_exp=grad_exp, grad_man=grad_man, use_kahan=False)


def kahan_sum_gradients(model: Module, grad_exp: int = 8, grad_man: int = 23) -> None:
    sum_gradients(model, use_APS=False, grad_exp=grad_exp, grad_man=grad_man, use_kahan=True)


def broadcast_params(model: Module, src: int = 0) -> None:
    """Broadcast every state_dict entry from ``src`` (dist_util.py:92-94).  Float tensors are
    coalesced into one flat broadcast instead of one collective per entry."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    tensors = [t for t in model.state_dict().values() if isinstance(t, torch.Tensor)]
    by_dtype: Dict[torch.dtype, list] = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for _dt, ts in by_dtype.items():
        flat = torch.cat([t.detach().reshape(-1) for t in ts])
        dist.broadcast(flat, src)
        off = 0
        with torch.no_grad():
            for t in ts:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n


class DistModule(Module):
    """Wrapper that synchronises the wrapped module's state from rank 0 at construction
    (dist_util.py:8-19).  Optionally owns the arena + reducer: ``DistModule(m, precision=cfg)``
    re-homes parameters/gradients into flat arenas and exposes ``sum_gradients()``."""

    def __init__(self, module: Module, precision: Optional[PrecisionConfig] = None):
        super().__init__()
        self.module = module
        broadcast_params(self.module)
        self.precision = precision
        self.arena: Optional[ModelArena] = None
        self.reducer: Optional[GradReducer] = None
        if precision is not None:
            self.arena = ModelArena(list(module.parameters()))
            module._cpd_arena = self.arena
            self.reducer = GradReducer(self.arena, precision)

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    def train(self, mode: bool = True):
        super().train(mode)
        self.module.train(mode)
        return self

    def sum_gradients(self) -> None:
        if self.reducer is None:
            raise RuntimeError("DistModule was built without a PrecisionConfig")
        self.arena.adopt_grads()
        self.reducer.reduce()
