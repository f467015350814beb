"""Fused all-reduce on real NVLink-connected GPUs (SURVEY §4 'Distributed, 2/4/8 GPUs'): every
variant == the gather-mode oracle bitwise on every rank, all ranks bit-identical, plus a
back-to-back launch stress.  Runs tools/multigpu_check.py under torchrun; skipped with < 2 GPUs."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
def test_fused_allreduce_matches_oracle_on_all_ranks(tmp_path):
    n = min(torch.cuda.device_count(), 8)
    out = tmp_path / "mg.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tools", "multigpu_check.py"),
           "--sections", "topo,correct,stress,fault", "--stress-iters", "300", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=850)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    assert res["correct_all_pass"] and res["stress_all_pass"] and res["fault_all_pass"], res
