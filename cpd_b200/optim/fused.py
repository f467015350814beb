"""Fused multi-tensor optimizer on fp32 master weights (LARS / momentum-SGD / Nesterov).

Reference: example/ResNet18/tools/mix.py:53-63 (fp32 master copies), :292-314 (route grads
to masters, per-tensor LARS python loop or torch SGD, copy masters back) — ~12 launches per
tensor per step.  ``FusedSGD`` does the whole model in ONE launch of lars_sm100.cu over
the flat arenas (norms -> trust ratio -> momentum/weight-decay update -> master->model
copy), reads the learning rate from device memory (CUDA-graph replay safe) and supports
per-tensor weight-decay multipliers (ResNet-50 recipe: no decay on BN, main.py:123-131).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from ..ops import native
from ..parallel.arena import ModelArena


class FusedSGD:
    """SGD / Nesterov / LARS on fp32 master weights in ONE multi-tensor launch per step (mix.py:292-314: master
    copy, ``local_lr = 0.001 * |w| / (|g| + wd * |w|)``, momentum buffer, model <- master), learning rate read from
    device memory so a captured CUDA graph can replay with a new value."""
    def __init__(self, arena: ModelArena, lr: float = 0.1, momentum: float = 0.9, weight_decay: float = 0.0,
                 use_lars: bool = False, lars_coeff: float = 0.001, lars_eps: float = 0.0, nesterov: bool = False,
                 wd_mult: Optional[Sequence[float]] = None, master_weights: bool = True):
        if arena.model_flat is None:
            raise ValueError("FusedSGD needs an arena that owns the parameters (rehome_params=True)")
        self.arena = arena
        self.momentum, self.weight_decay = float(momentum), float(weight_decay)
        self.use_lars, self.lars_coeff, self.lars_eps, self.nesterov = bool(use_lars), float(lars_coeff), float(lars_eps), bool(nesterov)
        dev = arena.device
        self.table = arena.table()
        # fp32 master copy (vestigial when the model is fp32 too, but part of the reference recipe);
        # without it the update is applied to the model arena directly
        self.master = arena.model_flat.clone() if master_weights else arena.model_flat
        self.model_out = arena.model_flat if master_weights else None
        self.mom = torch.zeros_like(arena.model_flat)
        self.lr_dev = torch.tensor([float(lr)], dtype=torch.float32, device=dev)
        self._lr_pinned = torch.tensor([float(lr)], dtype=torch.float32).pin_memory() if dev.type == "cuda" else None
        self.lr = float(lr)
        self.wd_mult = None
        if wd_mult is not None:
            if len(wd_mult) != arena.layout.T:
                raise ValueError("wd_mult needs one entry per parameter tensor")
            self.wd_mult = torch.tensor(list(wd_mult), dtype=torch.float32, device=dev)
        self.partials = None
        self.barrier = None
        if dev.type == "cuda":
            self.partials = torch.zeros(2 * max(self.table.n_chunks, 1), dtype=torch.float32, device=dev)
            self.barrier = torch.zeros(4, dtype=torch.int32, device=dev)
        self.steps = 0

    def set_lr(self, lr: float) -> None:
        """Update the device-resident learning rate (async pinned H2D copy; graph-safe)."""
        self.lr = float(lr)
        if self._lr_pinned is not None:
            self._lr_pinned[0] = self.lr
            self.lr_dev.copy_(self._lr_pinned, non_blocking=True)
        else:
            self.lr_dev[0] = self.lr

    def step(self) -> None:
        a = self.arena
        native.lars_step(self.master, self.model_out, a.grad_flat, self.mom, self.table.seg_off, lr=self.lr,
                         momentum=self.momentum, wd=self.weight_decay, coeff=self.lars_coeff, eps=self.lars_eps,
                         use_lars=self.use_lars, nesterov=self.nesterov, wd_mult=self.wd_mult, chunks=self.table.chunks,
                         tchunk_off=self.table.tchunk_off, lr_dev=self.lr_dev, partials=self.partials,
                         barrier_ws=self.barrier)
        self.steps += 1

    def zero_grad(self) -> None:
        self.arena.zero_grad()

    # -- checkpointing ---------------------------------------------------------------
    def state_dict(self) -> dict:
        return {"master": self.master.detach().cpu(), "momentum": self.mom.detach().cpu(), "lr": self.lr,
                "steps": self.steps, "hyper": dict(momentum=self.momentum, weight_decay=self.weight_decay,
                                                   use_lars=self.use_lars, lars_coeff=self.lars_coeff, nesterov=self.nesterov)}

    def resync_master(self) -> None:
        """Make the fp32 masters equal to the model arena (after model weights were loaded from a checkpoint that
        carries no compatible optimizer state: the step copies master -> model, so stale masters would silently
        revert the load)."""
        if self.model_out is not None:
            self.master.copy_(self.arena.model_flat)

    def load_state_dict(self, sd: dict) -> None:
        """Restore masters + momentum.  Foreign dicts are tolerated: a ``torch.optim.SGD`` state_dict (what the
        reference trainers save, mix.py:345-356 / main.py:261-269) contributes its momentum buffers and learning
        rate when the tensor shapes line up, the masters are then re-derived from the (already loaded) model."""
        if "master" in sd and "momentum" in sd:
            self.master.copy_(sd["master"])
            self.mom.copy_(sd["momentum"])
            if self.model_out is not None:
                self.model_out.copy_(self.master)
            self.set_lr(sd.get("lr", self.lr))
            self.steps = int(sd.get("steps", 0))
            return
        self.resync_master()
        state, groups = sd.get("state"), sd.get("param_groups")
        if isinstance(state, dict) and isinstance(groups, list):
            order = [i for g in groups for i in g.get("params", [])]
            views = [self.arena.layout.view(self.mom, t, like=p.data) for t, p in enumerate(self.arena.params)]
            if len(order) == len(views):
                for i, v in zip(order, views):
                    buf = (state.get(i) or {}).get("momentum_buffer")
                    if buf is not None and buf.numel() == v.numel():
                        v.copy_(buf.reshape(v.shape))
            if groups and "lr" in groups[0]:
                self.set_lr(float(groups[0]["lr"]))


def bn_bias_no_decay(named_params) -> list:
    """wd multipliers: 0 for BatchNorm parameters (names containing 'bn', as main.py:123-131 does), 1 otherwise."""
    return [0.0 if "bn" in n else 1.0 for n, p in named_params if p.requires_grad]
