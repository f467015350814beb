#!/usr/bin/env python
"""FCN-d8 semantic segmentation with low-precision gradient all-reduce — the reference's fourth workload
(FCN-R50-d8, 769x769 Cityscapes crops, 40 K iterations, trained through a fork of mmcv whose
``runner/hooks/optimizer.py`` calls the CPD gradient path; README.md:132-151).

Here the integration point is a library object: ``cpd_b200.parallel.APSOptimizerHook`` has mmcv's ``OptimizerHook``
interface, and the small iteration-based runner below drives hooks the way mmcv's ``IterBasedRunner`` does
(``after_train_iter`` with ``runner.outputs['loss']``), so the same hook drops into a real mmcv config
(``optimizer_config = dict(type='APSOptimizerHook', grad_exp=5, grad_man=2, use_APS=True)``).
Synthetic Cityscapes-shaped data (19 classes, 255 = ignore; no network for the real set), mmsegmentation's recipe:
SGD 0.01 / 0.9 / 5e-4, poly LR (power 0.9, min 1e-4), CE(main) + 0.4 CE(aux).

    python examples/FCN/train_fcn.py --arch fcn_r50_d8 --crop 769 --batch-size 2 --max-iters 40000 --use_APS
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/FCN/train_fcn.py --use_APS
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cpd_b200.models import create_model  # noqa: E402
from cpd_b200.parallel import APSOptimizerHook, DistModule, dist_init  # noqa: E402
from cpd_b200.utils.meters import AverageMeter  # noqa: E402


class SyntheticCityscapes(torch.utils.data.Dataset):
    """Blocky label maps (so that there is something to learn) + an image correlated with them; ~3 % ignore pixels."""
    def __init__(self, n, crop, num_classes=19, seed=0):
        self.n, self.crop, self.nc, self.seed = n, crop, num_classes, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        cells = max(2, self.crop // 32)
        coarse = torch.randint(0, self.nc, (1, 1, cells, cells), generator=g).float()
        label = F.interpolate(coarse, size=(self.crop, self.crop), mode="nearest")[0, 0].long()
        img = torch.randn(3, self.crop, self.crop, generator=g) * 0.5 + (label.float() / self.nc - 0.5).unsqueeze(0) * \
            torch.tensor([1.0, -1.0, 0.5]).view(3, 1, 1)
        label[torch.rand(self.crop, self.crop, generator=g) < 0.03] = 255
        return img, label


class PolyLRHook:
    def __init__(self, base_lr, max_iters, power=0.9, min_lr=1e-4):
        self.base_lr, self.max_iters, self.power, self.min_lr = base_lr, max_iters, power, min_lr

    def before_train_iter(self, runner):
        lr = (self.base_lr - self.min_lr) * (1 - runner.iter / self.max_iters) ** self.power + self.min_lr
        for g in runner.optimizer.param_groups:
            g["lr"] = lr


class IterRunner:
    """The part of mmcv's IterBasedRunner the hooks rely on: ``model``, ``optimizer``, ``iter``, ``outputs``."""
    def __init__(self, model, optimizer, hooks):
        self.model, self.optimizer, self.hooks, self.iter, self.outputs = model, optimizer, hooks, 0, {}

    def call(self, name):
        for h in self.hooks:
            fn = getattr(h, name, None)
            if fn is not None:
                fn(self)

    def train_iter(self, img, label):
        self.model.train()
        self.call("before_train_iter")
        out = self.model(img)
        main, aux = out if isinstance(out, tuple) else (out, None)
        loss = F.cross_entropy(main, label, ignore_index=255)
        if aux is not None:
            loss = loss + 0.4 * F.cross_entropy(aux, label, ignore_index=255)
        self.outputs = {"loss": loss, "num_samples": img.shape[0]}
        self.call("after_train_iter")                                    # APSOptimizerHook: backward, all-reduce, step
        self.iter += 1
        return float(loss.detach())


@torch.no_grad()
def mean_iou(model, loader, num_classes, device):
    model.eval()
    inter = torch.zeros(num_classes, dtype=torch.float64)
    union = torch.zeros(num_classes, dtype=torch.float64)
    for img, label in loader:
        pred = model(img.to(device)).argmax(1).cpu()
        keep = label != 255
        for c in range(num_classes):
            p, t = (pred == c) & keep, (label == c) & keep
            inter[c] += (p & t).sum()
            union[c] += (p | t).sum()
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        buf = torch.stack([inter, union]).to(device)
        torch.distributed.all_reduce(buf)
        inter, union = buf[0].cpu(), buf[1].cpu()
    valid = union > 0
    return float((inter[valid] / union[valid]).mean()) if valid.any() else 0.0


def main():
    ap = argparse.ArgumentParser(description="FCN-d8 / Cityscapes-shape segmentation with CPD gradients")
    ap.add_argument("--arch", default="fcn_r50_d8", choices=["fcn_r50_d8", "fcn_r18_d8"])
    ap.add_argument("--crop", type=int, default=769)
    ap.add_argument("--batch-size", type=int, default=2, help="per GPU (mmsegmentation: 2 x 4 GPUs)")
    ap.add_argument("--max-iters", type=int, default=40000)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--momentum", type=float, default=0.9)
    ap.add_argument("--weight-decay", type=float, default=5e-4)
    ap.add_argument("--grad_exp", type=int, default=5)
    ap.add_argument("--grad_man", type=int, default=2)
    ap.add_argument("--use_APS", "--use-APS", dest="use_APS", action="store_true")
    ap.add_argument("--use_kahan", "--use-kahan", dest="use_kahan", action="store_true")
    ap.add_argument("--grad-clip", type=float, default=0.0)
    ap.add_argument("--dataset-len", type=int, default=2975)
    ap.add_argument("--val-len", type=int, default=16)
    ap.add_argument("--val-freq", type=int, default=4000)
    ap.add_argument("--print-freq", type=int, default=50)
    ap.add_argument("--workers", type=int, default=2)
    ap.add_argument("--width", type=int, default=64, help="stem width (fcn_r18_d8 only; small values for smoke runs)")
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--fused-bn", action="store_true", help="BatchNormAct2d (fused BN + ReLU kernels) in the backbone")
    args = ap.parse_args()

    if args.cpu:
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
    rank, world = dist_init()
    device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() and not args.cpu else torch.device("cpu")
    torch.manual_seed(1234)
    kw = {"width": args.width} if args.arch == "fcn_r18_d8" else {}
    model = create_model(args.arch, num_classes=19, aux=True, fused_bn=args.fused_bn and device.type == "cuda", **kw).to(device)
    if device.type == "cuda":
        model = model.to(memory_format=torch.channels_last)
    model = DistModule(model)                                            # broadcasts rank 0's initial weights
    optimizer = torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay)
    hooks = [PolyLRHook(args.lr, args.max_iters),
             APSOptimizerHook(grad_exp=args.grad_exp, grad_man=args.grad_man, use_APS=args.use_APS, use_kahan=args.use_kahan,
                              grad_clip={"max_norm": args.grad_clip} if args.grad_clip > 0 else None)]
    runner = IterRunner(model, optimizer, hooks)

    train = SyntheticCityscapes(args.dataset_len, args.crop, seed=rank)
    val = SyntheticCityscapes(args.val_len, args.crop, seed=10_000 + rank)
    tl = torch.utils.data.DataLoader(train, batch_size=args.batch_size, shuffle=True, num_workers=args.workers, drop_last=True,
                                     pin_memory=device.type == "cuda")
    vl = torch.utils.data.DataLoader(val, batch_size=args.batch_size, num_workers=0)
    losses, t0 = AverageMeter(args.print_freq), time.time()
    if rank == 0:
        print(f"{args.arch}: {sum(p.numel() for p in model.parameters()) / 1e6:.1f} M parameters, crop {args.crop}, "
              f"batch {args.batch_size} x {world}, grad ({args.grad_exp},{args.grad_man}) APS={args.use_APS} Kahan={args.use_kahan}")
    while runner.iter < args.max_iters:
        for img, label in tl:
            img = img.to(device, non_blocking=True)
            if device.type == "cuda":
                img = img.contiguous(memory_format=torch.channels_last)
            loss = runner.train_iter(img, label.to(device, non_blocking=True))
            losses.update(loss)
            if rank == 0 and (runner.iter % args.print_freq == 0 or runner.iter == args.max_iters):
                dt = (time.time() - t0) / runner.iter
                print(f"Iter [{runner.iter}/{args.max_iters}]\tlr {optimizer.param_groups[0]['lr']:.5f}\tloss {losses.val:.4f} ({losses.avg:.4f})\t"
                      f"{dt:.3f} s/iter\teta {dt * (args.max_iters - runner.iter) / 3600:.2f} h", flush=True)
            if runner.iter % args.val_freq == 0 or runner.iter == args.max_iters:
                miou = mean_iou(model, vl, 19, device)
                if rank == 0:
                    print(f" * Iter {runner.iter}: mIoU {100 * miou:.2f}", flush=True)
            if runner.iter >= args.max_iters:
                break


if __name__ == "__main__":
    main()
