"""Rank-0 logging: table printer, TSV, structured JSONL, optional TensorBoard
(reference: DavidNet/utils.py:44-56, dawn.py:37-47, mix.py:168-171,323-343)."""
from __future__ import annotations

import json
import os
import time
from typing import Optional


class TableLogger:
    """Fixed-width column printer of the DavidNet recipe (utils.py:44-56): header on the first ``append``."""
    def __init__(self, rank: int = 0):
        self.rank, self.keys = rank, None

    def append(self, output: dict):
        if self.rank != 0:
            return
        if self.keys is None:
            self.keys = list(output.keys())
            print(*(f"{k:>12s}" for k in self.keys))
        filtered = [output[k] for k in self.keys]
        print(*(f"{v:12.4f}" if isinstance(v, float) else f"{v:12}" for v in filtered))


class TSVLogger:
    """Collects (epoch, hours, top1Accuracy) rows and prints them tab-separated (dawn.py:37-47, DAWNBench format)."""
    def __init__(self):
        self.log = ["epoch\thours\ttop1Accuracy"]

    def append(self, output: dict):
        self.log.append(f"{output['epoch']}\t{output['total time'] / 3600:.8f}\t{output.get('test acc', 0) * 100:.2f}")

    def __str__(self):
        return "\n".join(self.log)


class JsonlLogger:
    """One JSON object per line: step, loss, lr, device-timed img/s, comm stats ..."""

    def __init__(self, path: Optional[str], rank: int = 0):
        self.f = None
        if path and rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            self.f = open(path, "a", buffering=1)

    def log(self, **kv):
        if self.f:
            kv.setdefault("time", time.time())
            self.f.write(json.dumps(kv) + "\n")

    def close(self):
        if self.f:
            self.f.close()
            self.f = None


class ScalarWriter:
    """TensorBoard scalars if a writer implementation is importable, else a JSONL file with the
    same (tag, value, step) triples.  (The reference hard-requires tensorboardX, mix.py:16.)"""

    def __init__(self, logdir: str, rank: int = 0):
        self.w, self.j = None, None
        if rank != 0:
            return
        try:
            from torch.utils.tensorboard import SummaryWriter  # needs the `tensorboard` package
            self.w = SummaryWriter(logdir)
        except Exception:  # noqa: BLE001
            self.j = JsonlLogger(os.path.join(logdir, "scalars.jsonl"))

    def add_scalar(self, tag: str, value, step: int):
        if self.w:
            self.w.add_scalar(tag, value, step)
        elif self.j:
            self.j.log(tag=tag, value=float(value), step=int(step))

    def close(self):
        if self.w:
            self.w.close()
        if self.j:
            self.j.close()
