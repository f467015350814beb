#!/usr/bin/env python
"""Bring-up bisection: each stage is run in its own process (group) so a native crash is attributed."""
import faulthandler
import os
import sys

faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
stage = sys.argv[1]
import torch  # noqa: E402

if stage == "topo_single":
    from cpd_b200.ops import native
    torch.cuda.init()
    print("topology:", str(native.topology())[:400], flush=True)
elif stage == "nccl_only":
    import torch.distributed as dist
    import cpd_b200  # noqa: F401  (library loaded, not used)
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    t = torch.ones(4, device="cuda")
    dist.all_reduce(t)
    print("nccl ok", t.tolist(), flush=True)
    dist.destroy_process_group()
elif stage == "dist_init":
    from cpd_b200 import dist_init
    from cpd_b200.parallel import discover_topology
    r, w = dist_init()
    print("dist_init ok", r, w, discover_topology().summary(), flush=True)
elif stage.startswith("symm_"):
    import torch.distributed as dist
    from cpd_b200 import dist_init
    from cpd_b200.parallel.symm import SymmetricWorkspace
    r, w = dist_init()
    ws = SymmetricWorkspace(1 << 20, 64, backend=stage[5:])
    print(f"rank {r}: symm {stage[5:]} ok multicast={ws.has_multicast} ptrs={[hex(p) for p in ws.ptrs]} log={ws.log}", flush=True)
    dist.barrier()
elif stage.startswith("ar_"):
    # tiny fused all-reduce on the given backend
    import torch.distributed as dist
    from cpd_b200 import dist_init, PrecisionConfig
    from cpd_b200.parallel import GradReducer
    os.environ["CPD_B200_SYMM"] = stage[3:]
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from multigpu_check import FakeArena, fill, same_bits
    r, w = dist_init()
    dev = torch.device("cuda", torch.cuda.current_device())
    arena = FakeArena([10, 64, 4097, 70000], dev)
    fill(arena, r, 1)
    mine = arena.grad_flat.clone()
    orc = GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True), mode="gather")
    orc.reduce()
    want = arena.grad_flat.clone()
    for algo in ("oneshot", "twoshot", "nvls"):
        red = GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, algo=algo), mode="fused", timeout_cycles=4_000_000_000)
        arena.grad_flat.copy_(mine)
        red.reduce()
        torch.cuda.synchronize()
        try:
            red.check_status()
            st = "ok"
        except Exception as e:  # noqa: BLE001
            st = str(e)
        print(f"rank {r}: fused {algo} backend={red.symm.backend} mc={red.symm.has_multicast} match={same_bits(arena.grad_flat, want)} status={st}", flush=True)
    dist.barrier()
print("stage", stage, "done", flush=True)
