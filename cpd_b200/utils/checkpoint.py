"""Checkpoint save / load / auto-resume (reference: train_util.py:268-318, mix.py:345-356,
main.py:69-75,261-269,288-291).  One utility covers model, fp32 masters + momentum (which the
reference drops), LR / step / sampler position, RNG state and the precision config; rank 0
writes atomically, everyone else waits on a barrier; resume is actually wired."""
from __future__ import annotations

import glob
import os
import re
import shutil
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def _barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def save_checkpoint(state: dict, path: str, is_best: bool = False, rank: Optional[int] = None) -> None:
    rank = (dist.get_rank() if dist.is_available() and dist.is_initialized() else 0) if rank is None else rank
    if rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + ".tmp"
        torch.save(state, tmp)
        os.replace(tmp, path)
        if is_best:
            root, ext = os.path.splitext(path)
            shutil.copyfile(path, re.sub(r"(ckpt|checkpoint)[-_]?\d*$", r"\1", root) + "_best" + ext)
    _barrier()


def make_state(model, optimizer=None, step: int = 0, epoch: int = 0, best_prec1: float = 0.0, precision=None,
               extra: Optional[dict] = None) -> dict:
    sd = {"step": step, "epoch": epoch, "best_prec1": best_prec1,
          "state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
          "rng": {"torch": torch.get_rng_state(), "cuda": torch.cuda.get_rng_state_all() if torch.cuda.is_available() else None}}
    if optimizer is not None:
        sd["optimizer"] = optimizer.state_dict()
    if precision is not None:
        sd["precision"] = precision.to_dict()
    if extra:
        sd.update(extra)
    return sd


def load_state(path: str, model, optimizer=None, strict: bool = False, map_location="cpu") -> Tuple[float, int]:
    """Restore ``model`` (fixing up a ``module.`` prefix mismatch either way) and optionally the
    optimizer; returns ``(best_prec1, last_iter)`` like the reference's load_state."""
    if not os.path.isfile(path):
        raise FileNotFoundError(f"=> no checkpoint found at '{path}'")
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = ckpt.get("state_dict", ckpt.get("model", ckpt))
    own = set(model.state_dict().keys())
    if own and not (set(sd.keys()) & own):
        if all(k.startswith("module.") for k in sd):
            sd = {k[len("module."):]: v for k, v in sd.items()}
        elif all(("module." + k) in own for k in sd):
            sd = {"module." + k: v for k, v in sd.items()}
    with torch.no_grad():
        cur = model.state_dict()
        missing = [k for k in cur if k not in sd]
        for k, v in sd.items():
            if k in cur:
                cur[k].copy_(v)            # copy into the existing (arena-backed) storage
        if strict and missing:
            raise KeyError(f"missing keys in checkpoint: {missing[:5]}...")
    if optimizer is not None:
        # FusedSGD keeps fp32 masters that the step copies back into the model: re-derive them from the weights just
        # loaded, THEN let a compatible optimizer state (ours, or a torch.optim.SGD dict as the reference saves) refine
        if hasattr(optimizer, "resync_master"):
            optimizer.resync_master()
        if "optimizer" in ckpt:
            optimizer.load_state_dict(ckpt["optimizer"])
    if "rng" in ckpt and ckpt["rng"].get("torch") is not None:
        torch.set_rng_state(ckpt["rng"]["torch"])
    return float(ckpt.get("best_prec1", 0.0)), int(ckpt.get("step", ckpt.get("epoch", 0)))


def find_latest(pattern: str) -> Optional[str]:
    """Newest checkpoint matching a glob such as ``ckpt_*.pth`` or a ``checkpoint-{epoch}.pth.tar``
    format (main.py:69-75 scans epochs downward)."""
    if "{epoch}" in pattern:
        pattern = pattern.replace("{epoch}", "*")
    best, best_n = None, -1
    for p in glob.glob(pattern):
        nums = re.findall(r"\d+", os.path.basename(p))
        n = int(nums[-1]) if nums else 0
        if n > best_n and not p.endswith("_best.pth"):
            best, best_n = p, n
    return best
