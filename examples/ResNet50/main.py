#!/usr/bin/env python
"""ResNet-50 / ImageNet with custom-precision gradients (the workload of the reference's
example/ResNet50/main.py) on cpd_b200: 90 epochs, batch 32 per GPU, Nesterov SGD peak lr 3.2 with
5 warm-up epochs and x0.1 at 30/60/80, no weight decay on BN, emulate-node sub-batches, per-epoch
rank-0 checkpoints with auto-resume.  Synthetic ImageNet-shape data when --data is not given.

  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/ResNet50/main.py --grad_exp 5 --grad_man 2 --use-APS
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cpd_b200 import PrecisionConfig, dist_init  # noqa: E402
from cpd_b200.config import add_precision_args  # noqa: E402
from cpd_b200.data import DistributedSampler, imagenet_shape  # noqa: E402
from cpd_b200.models import create_model  # noqa: E402
from cpd_b200.optim import WarmupStepLR, bn_bias_no_decay  # noqa: E402
from cpd_b200.train import TrainEngine  # noqa: E402
from cpd_b200.train.loop import LoopConfig, run  # noqa: E402
from cpd_b200.utils import find_latest  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description="ResNet-50/ImageNet custom-precision training")
    ap.add_argument("--log-dir", default="./logs")
    ap.add_argument("--checkpoint-format", default="./checkpoint-{epoch}.pth.tar")
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--val-batch-size", type=int, default=32)
    ap.add_argument("--epochs", type=int, default=90)
    ap.add_argument("--base-lr", type=float, default=0.0125)
    ap.add_argument("--warmup-epochs", type=float, default=5)
    ap.add_argument("--momentum", type=float, default=0.9)
    ap.add_argument("--wd", type=float, default=1e-4)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--no-cuda", action="store_true", help="run on the CPU backend (C++ CPU ops, Gloo)")
    ap.add_argument("--data", default="", help="ImageFolder root with train/ and val/; synthetic if empty")
    ap.add_argument("--dataset-len", type=int, default=4096)
    ap.add_argument("--max-iter", type=int, default=0)
    ap.add_argument("--val-len", type=int, default=256)
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--arch", default="resnet50")
    add_precision_args(ap, grad_exp=5, grad_man=2)
    args = ap.parse_args()
    rank, world = dist_init() if ("RANK" in os.environ or "SLURM_PROCID" in os.environ) else (0, 1)
    torch.manual_seed(args.seed)
    torch.backends.cudnn.benchmark = True
    cfg = PrecisionConfig.from_args(args)
    model = create_model(args.arch)
    engine = TrainEngine(model, cfg, batch_size=args.batch_size, input_shape=(3, 224, 224), lr=args.base_lr, momentum=args.momentum,
                         weight_decay=args.wd, nesterov=True, wd_mult=bn_bias_no_decay(model.named_parameters()), fold_single=True,
                         device=torch.device("cpu") if args.no_cuda else None)
    if args.data:
        import torchvision.datasets as D
        import torchvision.transforms as T
        norm = T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
        train_ds = D.ImageFolder(os.path.join(args.data, "train"), T.Compose([T.RandomResizedCrop(224), T.RandomHorizontalFlip(), T.ToTensor(), norm]))
        val_ds = D.ImageFolder(os.path.join(args.data, "val"), T.Compose([T.Resize(256), T.CenterCrop(224), T.ToTensor(), norm]))
    else:
        train_ds, val_ds = imagenet_shape(args.dataset_len), imagenet_shape(args.val_len, seed=1)
    E = cfg.emulate_node
    it_per_epoch = len(train_ds) // (world * args.batch_size * E)
    peak = args.base_lr * world * E * (args.batch_size / 32.0) * 32        # main.py: lr scales with the global batch (3.2 at 8 x 32)
    sched = WarmupStepLR(0.1, max(peak, 0.1), int(args.warmup_epochs * it_per_epoch), [30 * it_per_epoch, 60 * it_per_epoch, 80 * it_per_epoch])
    sampler = DistributedSampler(train_ds, world_size=world, rank=rank)

    def batches():
        for epoch in range(args.epochs):
            sampler.set_epoch(epoch)
            yield from torch.utils.data.DataLoader(train_ds, batch_size=args.batch_size * E, sampler=sampler, num_workers=args.workers,
                                                   pin_memory=True, drop_last=True)

    def val_batches():
        return torch.utils.data.DataLoader(val_ds, batch_size=args.val_batch_size, pin_memory=True,
                                           sampler=DistributedSampler(val_ds, world_size=world, rank=rank, round_up=False, shuffle=False))

    save_dir = os.path.dirname(os.path.abspath(args.checkpoint_format)) or "."
    lc = LoopConfig(max_iter=args.max_iter or it_per_epoch * args.epochs, lr_fn=sched, print_freq=50, val_freq=it_per_epoch, profile_trace=args.profile_trace or None,
                    save_path=save_dir, log_jsonl=os.path.join(args.log_dir, "train.jsonl"), resume="auto", arch="resnet50")
    run(engine, batches(), lc, val_batches)


if __name__ == "__main__":
    main()
