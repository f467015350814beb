"""Flat fp32 arenas and chunk tables for the multi-tensor kernels.

The reference walks ``model.parameters()`` in python and launches tens of kernels per
tensor (dist_util.py:22-89, mix.py:251-314).  Here every per-parameter quantity
(gradient, model weight, master weight, momentum, emulate_node slots, wire buffers)
lives in one flat buffer with a shared layout: tensor ``t`` occupies
``[seg_off[t], seg_off[t+1])`` where every start is a multiple of 16 elements
(64 B) and the padding is kept at zero, so kernels only ever issue full 128-bit
vectors and a *chunk* (<= 4096 elements of ONE tensor) is the scheduling unit.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from ..ops import native

ALIGN = native.ALIGN_ELEMS
CHUNK = native.CHUNK_ELEMS


def _pad(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


@dataclass
class ChunkTable:
    """Device-side schedule for tensors [t0, t1) of a layout, offsets relative to ``base``."""
    t0: int
    t1: int
    base: int                    # element offset of the first tensor inside the arena
    numel: int                   # padded elements covered
    seg_off: torch.Tensor        # int64 CPU [T+1], relative to base
    chunks: torch.Tensor         # uint8 [n_chunks * 16] on the target device
    tchunk_off: torch.Tensor     # int32 [T+1] on the target device
    n_chunks: int

    @property
    def T(self) -> int:
        return self.t1 - self.t0


class ArenaLayout:
    """Offsets of a list of tensors inside one flat fp32 buffer (every tensor starts on a 16-element boundary,
    padding stays zero) and the chunk tables the multi-tensor kernels are scheduled by."""
    def __init__(self, numels: Sequence[int]):
        self.numels = [int(n) for n in numels]
        offs = [0]
        for n in self.numels:
            offs.append(offs[-1] + _pad(n))
        self.offsets = offs                      # len T+1, padded starts
        self.total = offs[-1]

    @property
    def T(self) -> int:
        return len(self.numels)

    def table(self, device, t0: int = 0, t1: Optional[int] = None) -> ChunkTable:
        t1 = self.T if t1 is None else t1
        base = self.offsets[t0]
        rel = np.asarray(self.offsets[t0:t1 + 1], dtype=np.int64) - base
        recs = []
        tco = [0]
        for i in range(t1 - t0):
            s, e = int(rel[i]), int(rel[i + 1])
            for c in range(s, e, CHUNK):
                recs.append((c, min(CHUNK, e - c), i))
            tco.append(len(recs))
        arr = np.array(recs, dtype=native.CHUNK_DTYPE) if recs else np.zeros(0, dtype=native.CHUNK_DTYPE)
        chunks = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
        return ChunkTable(t0=t0, t1=t1, base=base, numel=int(rel[-1]), seg_off=torch.from_numpy(rel.copy()),
                          chunks=chunks, tchunk_off=torch.tensor(tco, dtype=torch.int32, device=device),
                          n_chunks=len(recs))

    def new(self, device, zero: bool = True) -> torch.Tensor:
        return (torch.zeros if zero else torch.empty)(self.total, dtype=torch.float32, device=device)

    def view(self, flat: torch.Tensor, t: int, like: Optional[torch.Tensor] = None) -> torch.Tensor:
        """View of tensor ``t`` inside ``flat``; with ``like`` it takes its shape and (dense) strides."""
        n = self.numels[t]
        seg = flat[self.offsets[t]:self.offsets[t] + n]
        if like is None:
            return seg
        if like.is_contiguous() or like.numel() <= 1:
            return seg.view(like.shape)
        if _is_dense(like):
            return seg.as_strided(like.shape, like.stride())
        return seg.view(like.shape)


def _is_dense(t: torch.Tensor) -> bool:
    """True if the strides are a permutation of a contiguous layout (e.g. channels_last)."""
    dims = sorted(((st, sz) for st, sz in zip(t.stride(), t.shape) if sz > 1))
    expect = 1
    for st, sz in dims:
        if st != expect:
            return False
        expect *= sz
    return True


class ModelArena:
    """Re-homes a model's trainable parameters and their gradients into flat arenas.

    After construction ``p.data`` and ``p.grad`` of every trainable parameter are views
    into ``model_flat`` / ``grad_flat`` (memory format preserved), so the gradient sync
    and the optimizer are single launches over flat memory, and autograd accumulates
    straight into the bucket (like DDP's gradient_as_bucket_view).
    """

    def __init__(self, params: Sequence[torch.nn.Parameter], rehome_params: bool = True, bind_grads: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("ModelArena: no trainable parameters")
        dev = self.params[0].device
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("ModelArena: all parameters must be fp32 on one device")
        self.device = dev
        self.layout = ArenaLayout([p.numel() for p in self.params])
        self.model_flat = self.layout.new(dev) if rehome_params else None
        self.grad_flat = self.layout.new(dev)
        self.grad_views: List[torch.Tensor] = []
        with torch.no_grad():
            for t, p in enumerate(self.params):
                if rehome_params:
                    v = self.layout.view(self.model_flat, t, like=p.data)
                    v.copy_(p.data)
                    p.data = v
                g = self.layout.view(self.grad_flat, t, like=p.data)
                self.grad_views.append(g)
                if bind_grads:
                    if p.grad is not None:
                        g.copy_(p.grad)
                    p.grad = g
        self._tables = {}

    def adopt_grads(self) -> None:
        """Copy gradients that live outside the arena (fresh tensors made by autograd after a
        ``zero_grad(set_to_none=True)``) into it and re-point ``p.grad`` at the arena views."""
        src, dst = [], []
        for p, g in zip(self.params, self.grad_views):
            if p.grad is None:
                g.zero_()
                p.grad = g
            elif p.grad.data_ptr() != g.data_ptr():
                src.append(p.grad)
                dst.append(g)
        if src:
            with torch.no_grad():
                torch._foreach_copy_(dst, src)
            for p, g in zip(self.params, self.grad_views):
                p.grad = g

    def release_grads(self) -> None:
        """``p.grad = None`` for every parameter: the next backward then HANDS autograd's freshly written gradient
        tensors over instead of accumulating into the arena views (one read-modify-write kernel per parameter, plus the
        arena memset); ``gather_grads`` moves them into the arena with multi-tensor copies."""
        for p in self.params:
            p.grad = None

    def gather_grads(self, t0: int = 0, t1: Optional[int] = None) -> None:
        """Copy the gradients of parameters [t0, t1) from wherever autograd put them into the arena (ONE multi-tensor
        copy; parameters that received no gradient get zeros) and re-point ``p.grad`` at the arena views."""
        t1 = len(self.params) if t1 is None else t1
        src, dst = [], []
        for t in range(t0, t1):
            p, g = self.params[t], self.grad_views[t]
            if p.grad is None:
                g.zero_()
            elif p.grad.data_ptr() != g.data_ptr():
                src.append(p.grad)
                dst.append(g)
            p.grad = g
        if src:
            with torch.no_grad():
                torch._foreach_copy_(dst, src)

    def table(self, t0: int = 0, t1: Optional[int] = None) -> ChunkTable:
        key = (t0, t1)
        if key not in self._tables:
            self._tables[key] = self.layout.table(self.device, t0, t1)
        return self._tables[key]

    def bind_grads(self, flat: Optional[torch.Tensor] = None) -> None:
        """Point every ``p.grad`` at ``flat`` (default: the main gradient arena).  Used by
        emulate_node to let autograd write micro-batch k straight into slot k."""
        flat = self.grad_flat if flat is None else flat
        for t, p in enumerate(self.params):
            p.grad = self.layout.view(flat, t, like=p.data)

    def zero_grad(self, flat: Optional[torch.Tensor] = None) -> None:
        (self.grad_flat if flat is None else flat).zero_()

    def bucket_ranges(self, bucket_bytes: float) -> List[tuple]:
        """Contiguous tensor ranges of roughly ``bucket_bytes`` fp32 bytes, listed in the
        order gradients become ready (last parameters first).  0 -> one bucket."""
        T = self.layout.T
        if bucket_bytes <= 0:
            return [(0, T)]
        out, hi, acc = [], T, 0
        for t in range(T - 1, -1, -1):
            acc += self.layout.numels[t] * 4
            if acc >= bucket_bytes or (hi - t) >= 1024:
                out.append((t, hi))
                hi, acc = t, 0
        if hi > 0:
            out.append((0, hi))
        return out
