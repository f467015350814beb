"""Training engine: one call per optimizer step, everything on-device.

Replaces the reference trainers' inner loops (example/ResNet18/tools/mix.py:201-358,
example/ResNet50/main.py:141-212, example/DavidNet/utils.py:328-378) — python loops over
parameters with host syncs between hundreds of tiny kernels — by a fixed device-side
pipeline that is captured ONCE into a CUDA graph and replayed:

    H2D(inputs, pinned)  ->  [ E x (zero slot, forward, loss/(W*E), backward) ]
                         ->  emulate_node fold   (1 kernel, aps_fold_sm100.cu)
                         ->  gradient all-reduce (1 fused kernel per bucket, allreduce_fused_sm100a.cu)
                         ->  LARS / SGD on fp32 masters + master->model copy (1 kernel, lars_sm100.cu)
                         ->  D2H(loss)

Gradients, parameters, masters and momentum live in flat arenas (parallel/arena.py), the
learning rate is read from device memory, and the cross-GPU kernels keep their epochs on
the device, so graph replay needs no host-side bookkeeping.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from ..config import PrecisionConfig
from ..optim import FusedSGD
from ..parallel.api import broadcast_params
from ..parallel.arena import ModelArena
from ..parallel.emulate import EmulateNode
from ..parallel.reducer import GradReducer


_NVTX = os.environ.get("CPD_B200_NVTX", "0") == "1"      # NVTX ranges around fwd/bwd, gradient sync, optimizer


class TrainEngine:
    """One optimizer step per call, everything on the device (see the module docstring for the pipeline).

    ``train_step(images, labels, lr=None) -> loss`` is the user call (H2D of the batch, captured step, D2H of the loss);
    ``prefetch`` / ``train_step_pipelined`` hide the H2D copy behind the previous step; ``evaluate``, ``global_loss``,
    ``comm_stats``, ``state_dict`` / ``load_state_dict`` complete the trainer-facing surface."""
    def __init__(self, model: nn.Module, cfg: PrecisionConfig, *, batch_size: int, input_shape: Sequence[int] = (3, 32, 32),
                 lr: float = 0.1, momentum: float = 0.9, weight_decay: float = 1e-4, use_lars: bool = False,
                 nesterov: bool = False, wd_mult: Optional[Sequence[float]] = None, loss_reduction: str = "mean",
                 loss_scale: float = 1.0, channels_last: bool = True, cuda_graph: bool = True,
                 fold_single: bool = False, device: Optional[torch.device] = None, master_weights: bool = True,
                 lars_eps: float = 0.0, warmup_steps: int = 3, amp_dtype: Optional[torch.dtype] = None,
                 steal_grads: bool = True):
        self.cfg = cfg
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size() if self.distributed else 1
        self.rank = dist.get_rank() if self.distributed else 0
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = device
        self.is_cuda = device.type == "cuda"
        self.E = int(cfg.emulate_node)
        # steal_grads (E == 1): backward writes fresh gradient tensors (p.grad is None beforehand) and ONE multi-tensor copy
        # per bucket moves them into the arena, instead of a memset + one accumulate kernel per parameter
        self.steal_grads = bool(steal_grads) and self.E == 1
        self.batch_size = int(batch_size)
        self.loss_reduction, self.loss_scale = loss_reduction, float(loss_scale)
        # reduced-precision COMPUTE (the reference's dawn.py --half): autocast around forward + loss; parameters,
        # gradients, masters and the whole gradient path stay fp32, so the low-bit all-reduce semantics are untouched
        self.amp_dtype = amp_dtype

        model = model.to(device)
        self.channels_last = bool(channels_last and self.is_cuda)
        if self.channels_last:
            model = model.to(memory_format=torch.channels_last)
        self.model = model
        broadcast_params(model)                                   # DistModule semantics: rank 0's init everywhere
        self.arena = ModelArena(list(model.parameters()))
        model._cpd_arena = self.arena
        self.emulate = EmulateNode(self.arena, cfg, fold_single=fold_single)
        self.reducer = GradReducer(self.arena, cfg)
        self.reducer.gather_before_launch = bool(steal_grads) and self.E == 1
        self.optimizer = FusedSGD(self.arena, lr=lr, momentum=momentum, weight_decay=weight_decay, use_lars=use_lars,
                                  nesterov=nesterov, wd_mult=wd_mult, master_weights=master_weights, lars_eps=lars_eps)

        self.x = torch.zeros(self.E, self.batch_size, *input_shape, device=device)
        self.y = torch.zeros(self.E, self.batch_size, dtype=torch.long, device=device)
        self.loss_dev = torch.zeros(1, device=device)
        self.loss_host = torch.zeros(1).pin_memory() if self.is_cuda else torch.zeros(1)
        self.step_count = 0
        self.graph = None
        self._use_graph = bool(cuda_graph and self.is_cuda)
        self._warmup_left = int(warmup_steps) if self._use_graph else 0
        self._done_evt = torch.cuda.Event() if self.is_cuda else None
        self.kernel_launches_per_step = None
        self._x_stage = self._y_stage = self._copy_stream = self._staged_evt = self._consumed_evt = None
        self._stage_busy = self._consumed_recorded = False

    # ------------------------------------------------------------------ step body --
    def _loss(self, logits, target):
        red = "sum" if self.loss_reduction == "sum" else "mean"
        return F.cross_entropy(logits, target, reduction=red)

    def _step_body(self) -> None:
        from ..ops import native
        launches0 = native._C.launch_count()
        try:
            self._step_body_inner()
        finally:
            # cpd_b200 kernel launches of ONE step (counted by the native library); a captured graph replays exactly these
            self.kernel_launches_per_step = int(native._C.launch_count() - launches0)

    def _step_body_inner(self) -> None:
        nvtx = torch.cuda.nvtx if (self.is_cuda and _NVTX) else None
        self.loss_dev.zero_()
        denom = float(self.world * self.E)
        overlap = self.cfg.overlap and self.E == 1 and self.reducer._pending is not None      # hooks installed (CUDA)
        for k in range(self.E):
            if self.steal_grads:
                self.arena.release_grads()
            else:
                self.emulate.begin_micro(k)
            if nvtx:
                nvtx.range_push(f"fwd_bwd[{k}]")
            xin = self.x[k]
            if self.channels_last:
                xin = xin.contiguous(memory_format=torch.channels_last)
            with torch.autocast(device_type=self.device.type, dtype=self.amp_dtype or torch.bfloat16, enabled=self.amp_dtype is not None):
                out = self.model(xin)
            loss = self._loss(out.float(), self.y[k]) / denom
            if overlap:
                self.reducer.arm()
            (loss * self.loss_scale if self.loss_scale != 1.0 else loss).backward()
            self.loss_dev += loss.detach()
            if nvtx:
                nvtx.range_pop()
        if nvtx:
            nvtx.range_push("grad_sync")
        if self.steal_grads and not overlap:
            self.arena.gather_grads()
        self.emulate.fold()
        self.reducer.reduce()
        if nvtx:
            nvtx.range_pop()
            nvtx.range_push("optimizer")
        self.optimizer.step()
        if nvtx:
            nvtx.range_pop()

    def _capture(self) -> None:
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: DataLoader pin-memory / worker threads may legally call cudaHostAlloc etc. while we capture
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self._step_body()

    # ------------------------------------------------------------------- public --
    def set_lr(self, lr: float) -> None:
        self.optimizer.set_lr(lr)

    def load_batch(self, images: torch.Tensor, labels: torch.Tensor) -> None:
        """Host -> device copy of one step's inputs ([E*B, ...] or [E, B, ...]); pinned sources are async."""
        self.x.view(-1, *self.x.shape[2:]).copy_(images.view(-1, *self.x.shape[2:]), non_blocking=True)
        self.y.view(-1).copy_(labels.view(-1), non_blocking=True)

    # -- input pipelining: the H2D copy of step i+1 overlaps the compute of step i --------------------------------
    def prefetch(self, images: torch.Tensor, labels: torch.Tensor) -> None:
        """Start the host -> device copy of the NEXT step's inputs into a staging buffer on a copy stream (pinned
        sources make it asynchronous); `train_step_pipelined` / `begin_step` consume it."""
        if self._x_stage is None:
            self._x_stage, self._y_stage = torch.empty_like(self.x), torch.empty_like(self.y)
            if self.is_cuda:
                self._copy_stream = torch.cuda.Stream(device=self.device)
                self._staged_evt, self._consumed_evt = torch.cuda.Event(), torch.cuda.Event()
        if self.is_cuda:
            if self._consumed_recorded:
                self._copy_stream.wait_event(self._consumed_evt)          # the previous staged batch has been taken over
            with torch.cuda.stream(self._copy_stream):
                self._x_stage.view(-1, *self.x.shape[2:]).copy_(images.view(-1, *self.x.shape[2:]), non_blocking=True)
                self._y_stage.view(-1).copy_(labels.view(-1), non_blocking=True)
                self._staged_evt.record(self._copy_stream)
        else:
            self._x_stage.view(-1, *self.x.shape[2:]).copy_(images.view(-1, *self.x.shape[2:]))
            self._y_stage.view(-1).copy_(labels.view(-1))
        self._stage_busy = True

    def begin_step(self, lr: Optional[float] = None) -> None:
        """Take over the prefetched batch (device-to-device, a few us) and launch the step; returns immediately."""
        if not self._stage_busy:
            raise RuntimeError("begin_step() without a prefetched batch")
        if lr is not None:
            self.set_lr(lr)
        if self.is_cuda:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._staged_evt)
            self.x.copy_(self._x_stage, non_blocking=True)
            self.y.copy_(self._y_stage, non_blocking=True)
            self._consumed_evt.record(cur)
            self._consumed_recorded = True
        else:
            self.x.copy_(self._x_stage)
            self.y.copy_(self._y_stage)
        self.run_step()

    def end_step(self) -> float:
        """Device -> host read of the step's (local, already /W) loss."""
        if self.is_cuda:
            self.loss_host.copy_(self.loss_dev, non_blocking=True)
            self.reducer.status_async()                    # 4 more bytes next to the loss: a timed-out peer raises NOW
            self._done_evt.record()
            self._done_evt.synchronize()
            self.reducer.check_status_host()
        else:
            self.loss_host.copy_(self.loss_dev)
        return float(self.loss_host[0])

    def train_step_pipelined(self, next_images: Optional[torch.Tensor], next_labels: Optional[torch.Tensor],
                             lr: Optional[float] = None) -> float:
        """One step on the batch given to the previous `prefetch` / `train_step_pipelined` call while the next batch's
        H2D copy runs on the copy stream.  Prime with `prefetch(first batch)`; pass None, None on the last step."""
        self.begin_step(lr)
        self._stage_busy = False
        if next_images is not None:
            self.prefetch(next_images, next_labels)
        return self.end_step()

    def run_step(self) -> None:
        """One optimizer step on the inputs already in the static buffers (no host sync)."""
        if self._use_graph:
            if self._warmup_left > 0:
                s = torch.cuda.Stream(device=self.device)
                s.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(s):
                    self._step_body()
                torch.cuda.current_stream(self.device).wait_stream(s)
                self._warmup_left -= 1
            else:
                if self.graph is None:
                    self._capture()
                self.graph.replay()
        else:
            self._step_body()
        self.step_count += 1

    def train_step(self, images: torch.Tensor, labels: torch.Tensor, lr: Optional[float] = None) -> float:
        """The call a user makes: H2D the batch, one full step, D2H the (local, already /W) loss."""
        if lr is not None:
            self.set_lr(lr)
        self.load_batch(images, labels)
        self.run_step()
        if self.is_cuda:
            self.loss_host.copy_(self.loss_dev, non_blocking=True)
            self.reducer.status_async()                    # 4 more bytes next to the loss: a timed-out peer raises NOW
            self._done_evt.record()
            self._done_evt.synchronize()
            self.reducer.check_status_host()
        else:
            self.loss_host.copy_(self.loss_dev)
        return float(self.loss_host[0])

    @torch.no_grad()
    def evaluate(self, images: torch.Tensor, labels: torch.Tensor):
        was = self.model.training
        self.model.eval()
        x = images.to(self.device, non_blocking=True)
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        with torch.autocast(device_type=self.device.type, dtype=self.amp_dtype or torch.bfloat16, enabled=self.amp_dtype is not None):
            out = self.model(x)
        out = out.float()
        y = labels.to(self.device, non_blocking=True)
        loss = F.cross_entropy(out, y)
        self.model.train(was)
        return out, loss

    def global_loss(self) -> float:
        """Sum of the per-rank (already divided by W*E) losses = the global mean loss (logging)."""
        t = self.loss_dev.clone()
        if self.world > 1:
            dist.all_reduce(t)
        return float(t.item())

    def check(self) -> None:
        self.reducer.check_status()

    def comm_stats(self) -> dict:
        """Observability of the gradient path for the JSONL log: APS shift histogram of the last step,
        non-finite / all-zero tensor counters and, for the fused all-reduce, the device-stamped duration of the last launch.
        Host-synchronising (reads small device buffers) - call it at log points only."""
        out = {}
        if self.cfg.use_aps and getattr(self.reducer, "shifts", None):
            sh = torch.cat([s.flatten() for s in self.reducer.shifts]).cpu()
            if sh.numel():
                vals, counts = torch.unique(sh, return_counts=True)
                out["aps_shift_min"], out["aps_shift_max"] = int(sh.min()), int(sh.max())
                out["aps_shift_hist"] = {int(v): int(c) for v, c in zip(vals.tolist(), counts.tolist())}
        out.update(self.reducer.guard_counters())
        ph = self.reducer.phase_times_us()
        if ph.get("total") is not None:
            out["comm_us_last_launch"] = ph["total"]
        return out

    def state_dict(self) -> dict:
        return {"model": {k: v.detach().cpu() for k, v in self.model.state_dict().items()},
                "optimizer": self.optimizer.state_dict(), "step": self.step_count, "precision": self.cfg.to_dict()}

    def load_state_dict(self, sd: dict) -> None:
        with torch.no_grad():
            cur = self.model.state_dict()
            for k, v in sd["model"].items():
                cur[k].copy_(v)
        self.optimizer.resync_master()
        if sd.get("optimizer") is not None:
            self.optimizer.load_state_dict(sd["optimizer"])
        self.step_count = int(sd.get("step", 0))
