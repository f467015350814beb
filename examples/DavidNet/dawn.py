#!/usr/bin/env python
"""DavidNet / CIFAR-10 (the workload of the reference's example/DavidNet/dawn.py) on cpd_b200:
24 epochs, piecewise-linear LR [0,5,24] -> [0, 0.4*lr_scale, 0] divided by the batch size, Nesterov
SGD with wd = 5e-4 * batch, sum-reduced CE, Crop/FlipLR/Cutout.  Unlike the reference, --use_kahan
and --use_lars are available here too (BASELINE config: (5,2) + APS + Kahan + LARS).

  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/DavidNet/dawn.py --grad_exp 5 --grad_man 2 --use_APS
"""
import argparse
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cpd_b200 import PrecisionConfig, dist_init  # noqa: E402
from cpd_b200.config import add_precision_args  # noqa: E402
from cpd_b200.data import Crop, Cutout, FlipLR, Transform, cifar_shape, normalise, pad, transpose  # noqa: E402
from cpd_b200.models import create_model  # noqa: E402
from cpd_b200.optim import PiecewiseLinear, linear_warmup  # noqa: E402
from cpd_b200.train import TrainEngine  # noqa: E402
from cpd_b200.utils import DeviceTimer, TableLogger, TSVLogger  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dist", default=1, type=int)
    ap.add_argument("--epoch", default=24, type=int)
    ap.add_argument("--warm_up_epoch", default=5, type=int)
    ap.add_argument("-b", "--batch_size", default=512, type=int)
    ap.add_argument("--momentum", default=0.9, type=float)
    ap.add_argument("--workers", default=4, type=int)
    ap.add_argument("--lr_scale", default=1.0, type=float)
    ap.add_argument("--seed", default=0, type=int)
    ap.add_argument("--loss_scale", default=1, type=int)
    ap.add_argument("--half", default=0, type=int, help="1: fp16 forward/backward (autocast), as the reference's --half; gradients stay fp32")
    ap.add_argument("--bf16", action="store_true", help="bf16 forward/backward (autocast)")
    ap.add_argument("--data", default="")
    ap.add_argument("--dataset_len", default=50000, type=int)
    ap.add_argument("--cpu", action="store_true", help="run on the CPU backend (C++ CPU ops, Gloo)")
    add_precision_args(ap, grad_exp=8, grad_man=23)
    args = ap.parse_args()
    rank, world = dist_init() if (args.dist == 1 and "RANK" in os.environ or "SLURM_PROCID" in os.environ) else (0, 1)
    torch.manual_seed(args.seed)
    cfg = PrecisionConfig.from_args(args)
    B = args.batch_size
    engine = TrainEngine(create_model("davidnet"), cfg, batch_size=B, lr=0.0, momentum=args.momentum, weight_decay=5e-4 * B,
                         nesterov=not args.use_lars, use_lars=args.use_lars, loss_reduction="sum", loss_scale=args.loss_scale,
                         amp_dtype=torch.bfloat16 if args.bf16 else (torch.float16 if args.half else None),
                         device=torch.device("cpu") if args.cpu else None)
    if args.data:
        import torchvision
        raw = torchvision.datasets.CIFAR10(args.data, train=True)
        train = list(zip(transpose(normalise(pad(raw.data, 4))), raw.targets))
    else:
        ds = cifar_shape(args.dataset_len)
        train = list(zip(np.pad(ds.images.numpy(), [(0, 0), (0, 0), (4, 4), (4, 4)], mode="reflect"), ds.labels.tolist()))
    tset = Transform(train, [Crop(32, 32), FlipLR(), Cutout(8, 8)], seed=args.seed)
    sched = PiecewiseLinear([0, 5, args.epoch], [0, 0.4 * args.lr_scale, 0])
    per_rank = len(tset) // world
    steps_per_epoch = per_rank // B
    warm_up_iter = math.ceil(len(tset) * args.warm_up_epoch / (world * B))
    table, tsv, timer = TableLogger(rank), TSVLogger(), DeviceTimer()
    xb = torch.empty(B, 3, 32, 32).pin_memory() if torch.cuda.is_available() else torch.empty(B, 3, 32, 32)
    yb = torch.empty(B, dtype=torch.long).pin_memory() if torch.cuda.is_available() else torch.empty(B, dtype=torch.long)
    step, total_s = 0, 0.0
    for epoch in range(args.epoch):
        tset.set_random_choices()
        order = np.random.RandomState(args.seed + epoch).permutation(len(tset))[rank * per_rank:(rank + 1) * per_rank]
        timer.start()
        run_loss = 0.0
        for i in range(steps_per_epoch):
            idx = order[i * B:(i + 1) * B]
            for j, k in enumerate(idx):
                img, lab = tset[int(k)]
                xb[j] = torch.from_numpy(np.ascontiguousarray(img)); yb[j] = int(lab)
            lr = linear_warmup(sched(epoch + i / steps_per_epoch) / B, step, warm_up_iter)
            run_loss += engine.train_step(xb, yb, lr=lr)
            step += 1
        ms = timer.stop_ms()
        total_s += ms / 1e3
        summary = {"epoch": epoch + 1, "lr": lr * B, "train time": ms / 1e3, "train loss": run_loss * world / max(steps_per_epoch * B, 1),
                   "images/s": steps_per_epoch * B * world / (ms / 1e3), "total time": total_s}
        table.append(summary)
        tsv.append(summary)
    if rank == 0:
        print(tsv)


if __name__ == "__main__":
    main()
