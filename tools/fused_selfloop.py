#!/usr/bin/env python
"""Single-GPU 'self-loop' run of the fused all-reduce kernel (world_size 1: every peer pointer is
local memory).  Isolates the kernel's local passes (amax, pack, fold, unpack) and its barrier
protocol from NVLink, so it can be tuned and profiled with ncu on one GPU."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from cpd_b200 import PrecisionConfig  # noqa: E402
from cpd_b200.parallel import GradReducer  # noqa: E402
from kernel_microbench import time_ms, peaks  # noqa: E402
from multigpu_check import FakeArena, fill  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
hbm, _, src = peaks()
quick = "--quick" in sys.argv
dev = torch.device("cuda", 0)
from cpd_b200.models import create_model  # noqa: E402
layouts = {"resnet18 (11.2M, 62 tensors)": [p.numel() for p in create_model("resnet18_cifar").parameters()]}
if not quick:
    layouts["64Mi x 1 tensor"] = [1 << 26]
rows = []
for name, numels in layouts.items():
    arena = FakeArena(numels, dev)
    fill(arena, 0, 1, -8, -2)
    for algo in ("oneshot", "twoshot"):
        for kahan in (False, True):
            if quick and kahan:
                continue
            red = GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, use_kahan=kahan, algo=algo), mode="fused")
            ms = time_ms(red.reduce, iters=3 if quick else 10, warm=2)
            n = arena.layout.total
            rows.append({"layout": name, "algo": algo, "kahan": kahan, "us": round(ms * 1e3, 2), "elems": n,
                         "local_bytes_per_elem": 13, "frac_of_hbm_" + src: round(13.0 * n / ms / 1e6 / hbm, 3)})
            rows[-1]["phases_us"] = red.phase_times_us()
            print(json.dumps(rows[-1]), flush=True)
            del red
dist.destroy_process_group()
