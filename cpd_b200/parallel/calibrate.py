"""One-time calibration of the fused all-reduce's variant thresholds on the LIVE process group.

``Thresholds`` (parallel/reducer.py) derives the raw / one-shot / two-shot hand-over points from the discovered
topology (NVLink bandwidth x a flag-round latency constant).  ``calibrate()`` replaces the constants by measurements:
it times the three variants on this box, this world size, this wire width for a ladder of bucket sizes (device-timed,
max over ranks, fresh gradients every call), picks the crossover points and installs them as the default for every
``GradReducer`` created afterwards (``dist_init(calibrate=True)`` does this right after the process group is up).
A few hundred launches on one shared symmetric workspace: well under a second of GPU time.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist

from ..config import PrecisionConfig
from . import reducer as R
from .arena import ArenaLayout, ModelArena


class _Arena:
    """Just enough of ModelArena for GradReducer: a layout + a gradient arena."""

    def __init__(self, numels, device):
        self.layout = ArenaLayout(numels)
        self.device = device
        self.grad_flat = self.layout.new(device)
        self.model_flat = None
        self.params = []
        self._tables = {}

    def table(self, t0=0, t1=None):
        key = (t0, t1)
        if key not in self._tables:
            self._tables[key] = self.layout.table(self.device, t0, t1)
        return self._tables[key]

    bucket_ranges = ModelArena.bucket_ranges


def _time_us(fn, iters: int, group) -> float:
    for _ in range(8):
        fn()
    dist.barrier(group=group)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / iters], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def calibrate(grad_exp: int = 4, grad_man: int = 3, group=None, sizes: Optional[Sequence[int]] = None, iters: int = 30,
              install: bool = True, verbose: bool = False) -> R.Thresholds:
    """Measure raw vs packed one-shot vs two-shot / NVLS and return (and by default install) the thresholds."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2 or not torch.cuda.is_available():
        return R.default_thresholds(dist.get_world_size(group) if dist.is_initialized() else 1)
    from .symm import SymmetricWorkspace
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    sizes = list(sizes or [4096, 16384, 65536, 262144, 1048576, 4194304])          # elements
    base = R.default_thresholds(world)
    ws = SymmetricWorkspace(max(sizes) * 4, R.MAX_TENSORS_PER_BUCKET, group=group)
    big = "nvls" if ws.has_multicast else "twoshot"
    rows = []
    try:
        for n in sizes:
            arena = _Arena([n], dev)
            src = torch.randn(n, device=dev) * 1e-3
            row = {"elems": n}
            for algo in ("raw", "oneshot", big):
                red = R.GradReducer(arena, PrecisionConfig(grad_exp=grad_exp, grad_man=grad_man, use_aps=True, algo=algo),
                                    group=group, mode="fused", symm=ws)

                def call():
                    arena.grad_flat[:n].copy_(src)             # fresh (finite) gradients every call; the copy is in all three timings
                    red.reduce()
                row[algo] = _time_us(call, iters, group)
                red.close()
            rows.append(row)
            if verbose and dist.get_rank(group) == 0:
                print("[calibrate]", row, flush=True)
    finally:
        ws.close()
    # crossovers, scanned from the LARGEST size down (one noisy small-size sample must not end the scan): raw keeps the
    # buckets up to the largest size at which it still beats both packed variants; one-shot likewise against two-shot
    raw_max = 1024
    for r in reversed(rows):
        if r["raw"] <= min(r["oneshot"], r[big]) * 1.02:
            raw_max = r["elems"]
            break
    one_max = 64 * 1024
    wire_bytes = R.native.wire_bits(grad_exp, grad_man) // 8
    for r in reversed(rows):
        if r["oneshot"] <= r[big] * 1.02:
            one_max = r["elems"] * wire_bytes
            break
    th = R.Thresholds(world, base.link_gbs, base.multicast, raw_max_elems=raw_max,
                      oneshot_max_bytes=one_max, source="calibrated")
    th.rows = rows
    if install:
        R._CALIBRATED[world] = th
    return th
