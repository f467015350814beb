"""Shared training / validation loop for the three reference workloads.

Reference trainers: example/ResNet18/tools/mix.py (iteration-based, YAML config, validate every
val_freq, rank-0 checkpoints + TensorBoard), example/DavidNet/dawn.py (epoch-based, piecewise
linear LR, TSV log), example/ResNet50/main.py (epoch-based, auto-resume).  One loop serves the
iteration-based recipes on top of ``TrainEngine``; throughput is CUDA-event timed and reported
as the max over ranks; logging reductions run every ``print_freq`` steps instead of every step.
"""
from __future__ import annotations

import contextlib
import os
from dataclasses import dataclass
from typing import Callable, Iterable, Optional

import torch
import torch.distributed as dist

from ..utils import AverageMeter, DeviceTimer, JsonlLogger, ScalarWriter, accuracy, find_latest, load_state, make_state, save_checkpoint, trace
from .engine import TrainEngine


@dataclass
class LoopConfig:
    max_iter: int
    lr_fn: Callable[[int], float]
    print_freq: int = 50
    val_freq: int = 0
    save_path: Optional[str] = None
    log_jsonl: Optional[str] = None
    tensorboard_dir: Optional[str] = None
    resume: Optional[str] = None          # checkpoint path or "auto"
    arch: str = "model"
    profile_trace: Optional[str] = None   # chrome trace (torch.profiler) of `profile_steps` steps starting at `profile_start`
    profile_start: int = 10
    profile_steps: int = 3


def validate(engine: TrainEngine, val_batches: Iterable, rank: int, world: int):
    """All-reduced loss / top-1 / top-5 over the validation stream (mix.py:361-429)."""
    tot = torch.zeros(4, device=engine.device)
    for x, y in val_batches:
        out, loss = engine.evaluate(x, y)
        p1, p5 = accuracy(out, y.to(engine.device), topk=(1, min(5, out.shape[1])))
        n = float(x.shape[0])
        tot += torch.stack([loss * n, p1 * n, p5 * n, torch.tensor(n, device=engine.device)])
    if world > 1:
        dist.all_reduce(tot)
    n = max(float(tot[3]), 1.0)
    return float(tot[0]) / n, float(tot[1]) / n, float(tot[2]) / n


def run(engine: TrainEngine, train_batches: Iterable, cfg: LoopConfig, val_batches_fn: Optional[Callable[[], Iterable]] = None):
    rank, world = engine.rank, engine.world
    jl = JsonlLogger(cfg.log_jsonl, rank)
    tb = ScalarWriter(cfg.tensorboard_dir, rank) if cfg.tensorboard_dir else None
    start_iter, best_prec1 = 0, 0.0
    if cfg.resume:
        path = cfg.resume
        if path == "auto":
            path = find_latest(os.path.join(cfg.save_path, "ckpt_*.pth")) if cfg.save_path else None
        if path and os.path.isfile(path):
            best_prec1, start_iter = load_state(path, engine.model, engine.optimizer)
            if rank == 0:
                print(f"=> resumed from {path} at iteration {start_iter}")
    losses = AverageMeter(cfg.print_freq)
    timer = DeviceTimer()
    imgs_per_step = engine.batch_size * engine.E * world
    timer.start()
    window, step = 0, start_iter
    prof_stack = contextlib.ExitStack()
    for x, y in train_batches:
        if step >= cfg.max_iter:
            break
        step += 1
        if cfg.profile_trace and step == start_iter + cfg.profile_start:
            prof_stack.enter_context(trace(cfg.profile_trace, rank))
        lr = cfg.lr_fn(step)
        engine.train_step(x, y, lr=lr)
        if cfg.profile_trace and step == start_iter + cfg.profile_start + cfg.profile_steps - 1:
            if engine.is_cuda:
                torch.cuda.synchronize()
            prof_stack.close()
        window += 1
        if step == 1 or step % cfg.print_freq == 0:
            gl = engine.global_loss()
            losses.update(gl)
            ms = timer.stop_ms()
            ips = imgs_per_step * window / (ms / 1e3)
            stats = engine.comm_stats()
            if rank == 0:
                print(f"Iter: [{step}/{cfg.max_iter}]\tLoss {gl:.4f} ({losses.avg:.4f})\tLR {lr:.4f}\t"
                      f"{ips:,.0f} img/s (device-timed, max over ranks)", flush=True)
                jl.log(step=step, loss=gl, lr=lr, images_per_sec=ips, ms_per_step=ms / window, **stats)
                if tb:
                    tb.add_scalar("loss_train", losses.avg, step)
                    tb.add_scalar("lr", lr, step)
            engine.check()
            window = 0
            timer.start()
        if cfg.val_freq and val_batches_fn is not None and (step % cfg.val_freq == 0 or step == cfg.max_iter):
            vl, p1, p5 = validate(engine, val_batches_fn(), rank, world)
            if rank == 0:
                print(f" * All Loss {vl:.4f} Prec@1 {p1:.3f} Prec@5 {p5:.3f}", flush=True)
                jl.log(step=step, loss_val=vl, acc1_val=p1, acc5_val=p5)
                if tb:
                    tb.add_scalar("loss_val", vl, step)
                    tb.add_scalar("acc1_val", p1, step)
                    tb.add_scalar("acc5_val", p5, step)
            is_best = p1 > best_prec1
            best_prec1 = max(best_prec1, p1)
            if cfg.save_path:
                save_checkpoint(make_state(engine.model, engine.optimizer, step=step, best_prec1=best_prec1, precision=engine.cfg,
                                           extra={"arch": cfg.arch}), os.path.join(cfg.save_path, f"ckpt_{step}.pth"), is_best, rank)
            timer.start()
            window = 0
    prof_stack.close()
    jl.close()
    if tb:
        tb.close()
    return step, best_prec1
