#!/usr/bin/env python
"""ResNet-18 / CIFAR-10 with custom-precision gradients (the workload of the reference's
example/ResNet18/tools/mix.py), on the cpd_b200 engine.

  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/ResNet18/mix.py --dist --grad_exp 4 --grad_man 3 --use_APS
  python examples/ResNet18/mix.py --emulate_node 8 --grad_exp 4 --grad_man 3 --use_APS        # 1 GPU plays 8 nodes
Flags keep the reference's names (mix.py:29-42); both spellings are accepted.
"""
import argparse
import math
import os
import sys

import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cpd_b200 import PrecisionConfig, dist_init  # noqa: E402
from cpd_b200.config import add_precision_args  # noqa: E402
from cpd_b200.data import DistributedGivenIterationSampler, DistributedSampler, cifar_shape  # noqa: E402
from cpd_b200.models import create_model  # noqa: E402
from cpd_b200.optim import WarmupStepLR  # noqa: E402
from cpd_b200.train import TrainEngine  # noqa: E402
from cpd_b200.train.loop import LoopConfig, run  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description="ResNet-18/CIFAR custom-precision training")
    ap.add_argument("--config", default=os.path.join(os.path.dirname(__file__), "configs/res18_cifar.yaml"))
    ap.add_argument("--dist", action="store_true")
    ap.add_argument("--load-path", "--load_path", dest="load_path", default="")
    ap.add_argument("--resume-opt", "--resume_opt", dest="resume_opt", action="store_true")
    ap.add_argument("-e", "--evaluate", action="store_true")
    ap.add_argument("--data", default="", help="CIFAR-10 root (torchvision layout); synthetic data of the same shape if empty")
    ap.add_argument("--max-iter", dest="max_iter_override", type=int, default=0)
    # command-line overrides of the YAML's `common:` keys (the reference can only edit the file, mix.py:69-72)
    for key, typ in (("arch", str), ("batch_size", int), ("workers", int), ("dataset_len", int), ("val_freq", int),
                     ("print_freq", int), ("save_path", str), ("max_epoch", int)):
        ap.add_argument("--" + key, "--" + key.replace("_", "-"), dest="cli_" + key, type=typ, default=None)
    ap.add_argument("--val-len", dest="val_len", type=int, default=2048)
    ap.add_argument("--cpu", action="store_true", help="run on the CPU backend (C++ CPU ops, Gloo)")
    add_precision_args(ap, grad_exp=5, grad_man=2)
    args = ap.parse_args()
    with open(args.config) as f:
        for k, v in yaml.safe_load(f)["common"].items():
            cli = getattr(args, "cli_" + k, None)
            setattr(args, k, v if cli is None else cli)
    rank, world = dist_init() if args.dist else (0, 1)
    cfg = PrecisionConfig.from_args(args)
    engine = TrainEngine(create_model(args.arch), cfg, batch_size=args.batch_size, lr=args.base_lr, momentum=args.momentum,
                         weight_decay=args.weight_decay, use_lars=args.use_lars,
                         device=torch.device("cpu") if args.cpu else None)
    if args.data:
        import torchvision
        import torchvision.transforms as T
        norm = T.Normalize((0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010))
        train_ds = torchvision.datasets.CIFAR10(args.data, train=True, transform=T.Compose(
            [T.RandomCrop(32, padding=4), T.RandomHorizontalFlip(), T.ToTensor(), norm]))
        val_ds = torchvision.datasets.CIFAR10(args.data, train=False, transform=T.Compose([T.ToTensor(), norm]))
    else:
        train_ds, val_ds = cifar_shape(args.dataset_len), cifar_shape(args.val_len, seed=1)
    E = cfg.emulate_node
    it_per_epoch = len(train_ds) / (world * args.batch_size * E)
    max_iter = args.max_iter_override or math.ceil(it_per_epoch * args.max_epoch)                  # mix.py:131-132
    sched = WarmupStepLR(args.base_lr, args.peak_lr, int(args.warmup_epochs * it_per_epoch),
                         [int(s * it_per_epoch) for s in args.lr_steps], args.lr_mult)
    sampler = DistributedGivenIterationSampler(train_ds, max_iter, args.batch_size * E, world_size=world, rank=rank)
    loader = torch.utils.data.DataLoader(train_ds, batch_size=args.batch_size * E, shuffle=False, num_workers=args.workers,
                                         pin_memory=True, sampler=sampler, drop_last=True)
    def val_batches():
        return torch.utils.data.DataLoader(val_ds, batch_size=args.batch_size, shuffle=False, num_workers=0, pin_memory=True,
                                           sampler=DistributedSampler(val_ds, world_size=world, rank=rank, round_up=False, shuffle=False))
    lc = LoopConfig(max_iter=max_iter, lr_fn=sched, print_freq=args.print_freq, val_freq=args.val_freq, save_path=args.save_path, profile_trace=args.profile_trace or None,
                    log_jsonl=os.path.join(args.save_path, "train.jsonl"), tensorboard_dir=os.path.join(args.save_path, "events"),
                    resume=(args.load_path or None), arch=args.arch)
    if args.evaluate:
        from cpd_b200.train.loop import validate
        if args.load_path:
            from cpd_b200.utils import load_state
            load_state(args.load_path, engine.model, engine.optimizer if args.resume_opt else None)
        print(" * All Loss {:.4f} Prec@1 {:.3f} Prec@5 {:.3f}".format(*validate(engine, val_batches(), rank, world)))
        return
    run(engine, loader, lc, val_batches)


if __name__ == "__main__":
    main()
