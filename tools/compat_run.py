#!/usr/bin/env python
"""Library-swap check: the reference's training recipes (its own example models, its per-parameter loops, restated
line by line in baseline/ref_harness.py) run on `--lib reference` (the unmodified drcut/CPD install) or on `--lib ours`
(this repo's import-compatible CPDtorch facade).  Prints the per-step global losses as one JSON line, so the two
libraries can be compared on the same seed:

    python tools/compat_run.py --lib reference --model resnet18 --steps 6 > ref.json
    python tools/compat_run.py --lib ours      --model resnet18 --steps 6 > ours.json
    (multi-GPU: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/compat_run.py ...)

The gradient path is bit-identical by construction; the convolutions' backward is not deterministic in cuDNN, so the
loss curves agree to a few 1e-3 relative rather than bitwise."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", choices=["reference", "ours"], default="ours")
    ap.add_argument("--model", choices=["resnet18", "resnet50", "davidnet"], default="resnet18")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--grad_exp", type=int, default=4)
    ap.add_argument("--grad_man", type=int, default=3)
    ap.add_argument("--no-aps", action="store_true")
    ap.add_argument("--use_kahan", action="store_true")
    ap.add_argument("--use_lars", action="store_true")
    args = ap.parse_args()
    import ref_harness
    if args.lib == "reference":
        ok, why = ref_harness.available()
        if not ok:
            print(json.dumps({"lib": "reference", "unavailable": why}))
            return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    ref_harness.setup_env(rank, world, lib=args.lib)
    import torch
    torch.manual_seed(0)
    torch.backends.cudnn.deterministic = True
    batch = args.batch or {"resnet18": 128, "resnet50": 16, "davidnet": 128}[args.model]
    runner = ref_harness.RefRunner(args.model, batch, args.grad_exp, args.grad_man, not args.no_aps, args.use_kahan, args.use_lars)
    import CPDtorch
    losses = [runner.step() for _ in range(args.steps)]
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"lib": args.lib, "CPDtorch": os.path.dirname(CPDtorch.__file__), "model": args.model, "world": world,
                          "losses": [round(float(v), 6) for v in losses]}), flush=True)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
