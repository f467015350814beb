#!/usr/bin/env python
"""Run the all-reduce threshold calibration on the live group and print what it measured (torchrun entry point)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpd_b200 import dist_init  # noqa: E402
from cpd_b200.parallel.calibrate import calibrate  # noqa: E402
from cpd_b200.parallel.reducer import Thresholds, default_thresholds  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world = dist_init(verbose=True)
topo_th = default_thresholds(world)
th = calibrate(verbose=True)
if rank == 0:
    print("[calibrate] topology-derived thresholds:", json.dumps(topo_th.to_dict()))
    print("[calibrate] measured thresholds:", json.dumps(th.to_dict()))
    print("[calibrate] rows:", json.dumps(getattr(th, "rows", [])))
dist.barrier()
dist.destroy_process_group()
