"""``emulate_node``: one GPU plays E data-parallel nodes.

Reference: example/ResNet18/tools/mix.py:215-285 and example/ResNet50/main.py:156-202 —
every micro-batch's gradients are cloned into per-parameter python lists, then per
parameter: E x {log2, ceil, .cpu()} host syncs, E x {scale, quantize, copy}, E x
{add, quantize}, unscale.  Here the E micro-gradients are written by autograd directly
into E slot arenas and ONE kernel (aps_fold_sm100.cu) does amax-over-slots -> shift ->
slot-ordered quantised fold -> unscale for the whole model.
"""
from __future__ import annotations

from typing import List

import torch

from ..config import PrecisionConfig
from ..ops import native
from .arena import ModelArena


class EmulateNode:
    """One GPU plays ``cfg.emulate_node`` data-parallel nodes: autograd writes micro-batch k into slot arena k
    (``begin_micro``), ``fold`` reduces the slots with the reference's local low-precision fold in one kernel
    (mix.py:251-285, main.py:178-202).  Up to 64 slots."""
    def __init__(self, arena: ModelArena, cfg: PrecisionConfig, fold_single: bool = False):
        self.arena = arena
        self.cfg = cfg
        self.E = int(cfg.emulate_node)
        # mix.py:256-258 skips the fold when E == 1; main.py runs it anyway (fold_single=True)
        self.active = self.E > 1 or fold_single
        self.slots: List[torch.Tensor] = []
        self.table = None
        self.ws = None
        self.shifts = None
        if self.active:
            if self.E > 64:
                raise ValueError("emulate_node > 64 is not supported by the fold kernel")
            self.slots = [arena.layout.new(arena.device) for _ in range(self.E)]
            self.table = arena.table()
            self.shifts = torch.zeros(self.table.T, dtype=torch.int32, device=arena.device)
            if arena.device.type == "cuda":
                self.ws = torch.zeros(4 + 2 * max(self.table.T, 1), dtype=torch.int32, device=arena.device)

    def begin_micro(self, k: int) -> None:
        """Route the next backward's gradients into slot ``k`` (zeroed)."""
        if not self.active:
            if k == 0:
                self.arena.zero_grad()
            return
        self.slots[k].zero_()
        self.arena.bind_grads(self.slots[k])

    def fold(self) -> None:
        """slots -> arena.grad_flat with the reference's local low-precision fold
        (never Kahan, mix.py:278-282), then re-point ``p.grad`` at the main arena."""
        if not self.active:
            return
        cfg = self.cfg
        native.aps_fold(self.slots, self.arena.grad_flat, self.table.seg_off, cfg.grad_exp, cfg.grad_man,
                        cfg.use_aps, False, float(self.E), chunks=self.table.chunks, workspace=self.ws,
                        shifts_out=self.shifts, rounding=cfg.rounding)
        self.arena.bind_grads()
