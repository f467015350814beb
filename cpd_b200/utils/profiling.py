"""Optional tracing helpers.  The reference has no profiling hooks at all; here the training step
carries NVTX ranges (``CPD_B200_NVTX=1``: fwd_bwd[k] / grad_sync / optimizer, see train/engine.py) and any
region can be exported as a chrome trace with ``torch.profiler``.  Numbers taken while a profiler is attached are
never bench values - bench.py does not use this module."""
import contextlib
import os
from typing import Optional

import torch


@contextlib.contextmanager
def nvtx_range(name: str, enabled: Optional[bool] = None):
    """NVTX range when CUDA is available and tracing is on (env CPD_B200_NVTX=1 or enabled=True)."""
    on = (os.environ.get("CPD_B200_NVTX", "0") == "1") if enabled is None else enabled
    on = on and torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


@contextlib.contextmanager
def trace(path: Optional[str], rank: int = 0, record_shapes: bool = False, with_stack: bool = False):
    """``with trace("step.json"): engine.train_step(...)`` writes a chrome trace (CPU + CUDA activities) on rank 0.
    ``path`` None or empty disables it (zero overhead).  Yields the profiler (or None) so callers can ``prof.step()``."""
    if not path or rank != 0:
        yield None
        return
    acts = [torch.profiler.ProfilerActivity.CPU]
    if torch.cuda.is_available():
        acts.append(torch.profiler.ProfilerActivity.CUDA)
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    with torch.profiler.profile(activities=acts, record_shapes=record_shapes, with_stack=with_stack) as prof:
        yield prof
    prof.export_chrome_trace(path)


def kernel_table(prof, row_limit: int = 25) -> str:
    """Top kernels of a finished ``trace`` by device time (text table)."""
    key = "cuda_time_total" if torch.cuda.is_available() else "cpu_time_total"
    return prof.key_averages().table(sort_by=key, row_limit=row_limit)
