#!/usr/bin/env python
"""Fault injection for the fused all-reduce's bounded waits, on ONE GPU.

The kernel is told world_size = 2 while rank 1 does not exist: both ranks' pointers alias this GPU's region, so this
rank's own flag arrives and "rank 1"'s never does.  With a short timeout the kernel must (a) return instead of hanging,
(b) poison its outputs with NaN instead of folding stale peer data, (c) leave a status word that makes
``check_status()`` raise a RuntimeError naming the missing rank and the phase.  Multi-GPU flavour of the same check:
``CPD_B200_FAULT_DELAY_RANK=r`` (tests/test_multigpu.py)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from cpd_b200 import PrecisionConfig, _C  # noqa: E402
from cpd_b200.parallel import GradReducer  # noqa: E402
from multigpu_check import FakeArena, fill  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29544")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
dev = torch.device("cuda", 0)
out = []
for algo in ("raw", "oneshot", "twoshot"):
    arena = FakeArena([4096, 300, 64], dev)
    fill(arena, 0, 5, -6, 0)
    red = GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, algo=algo), mode="fused",
                      timeout_cycles=4_000_000)                      # ~2 ms
    s = red.symm
    base = s.ptrs[0]
    s.buffers = _C.CommBuffers(0, 2, [base + s.off_data] * 2, [base + s.off_gather] * 2, [base + s.off_mailbox] * 2,
                               [base + s.off_flags] * 2, 0, s.cap_bytes, s.max_tensors)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    red.reduce()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    msg = ""
    try:
        red.check_status()
    except RuntimeError as e:
        msg = str(e)
    nan_frac = float(torch.isnan(arena.grad_flat[:4096]).float().mean())
    out.append(dict(algo=algo, returned_ms=round(ms, 2), all_nan=nan_frac == 1.0, error=msg))
    del red
print("FAULT_TIMEOUT " + json.dumps(out), flush=True)
dist.destroy_process_group()
