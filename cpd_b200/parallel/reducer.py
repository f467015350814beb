"""Gradient reduction engine: the data-parallel low-precision all-reduce of CPD.

Reference: ``sum_gradients`` (CPDtorch/utils/dist_util.py:22-89) — python loops over
parameters, ~46 launches + 2 NCCL calls + host syncs per parameter, fp32 on the wire,
no overlap.  ``GradReducer`` reduces a whole ``ModelArena`` in a handful of launches:

  mode ``local``   world_size == 1: one fused APS-quantise kernel (aps_fold_sm100.cu)
  mode ``fused``   world_size  > 1 on NVLink-connected GPUs: ONE sm_100a kernel per
                   bucket that does amax exchange, scale, cast, pack, the rank-ordered
                   (Kahan) fold over peer memory and the unscale
                   (allreduce_fused_sm100a.cu; one-shot / two-shot / NVLS by size)
  mode ``gather``  fallback + test oracle: the reference algorithm — all_gather of
                   fp32 gradients (NCCL or Gloo) followed by one fold kernel over the
                   W gathered copies.  Bit-identical to ``fused`` by construction.

Buckets can be launched from autograd hooks on a dedicated comm stream so the
reduction overlaps the rest of backward (``overlap=True``).
"""
from __future__ import annotations

import os
import warnings
from typing import List, Optional

import torch
import torch.distributed as dist

from ..config import PrecisionConfig
from ..ops import native
from .arena import ChunkTable, ModelArena

_ALGO = {"auto": native.ALGO_AUTO, "oneshot": native.ALGO_ONESHOT, "twoshot": native.ALGO_TWOSHOT,
         "nvls": native.ALGO_NVLS, "raw": native.ALGO_ONESHOT_RAW}

# Variant thresholds, from the 8 x B200 sweep with fresh (finite) gradients per call (profiles/allreduce_trip18_n8_*):
#   fp32 payload      raw    one-shot   two-shot   NVLS    (us, (4,3) + APS, 8-bit wire)
#   1 KB / 4 KB      21.3 / 21.9   29.1 / 29.3   35.6    35.2
#   16 KB .. 1 MB    43.6 .. 51.5  30.4 .. 34.7  37.6 .. 39.9
#   4 MB             78.6    48.4       46.5       45.7
#   16 MB / 64 MB      -     103.6 / 293   69.7 / 167   67.8 / 166
# -> raw (one flag round, fp32 on the wire) up to 1 K elements, packed one-shot up to 512 KB of wire codes, NVLS /
#    two-shot above.  (The raw variant's jump between 4 KB and 16 KB is not understood yet - open item.)
# opt-in: phase B of the fused kernel packs (4,3)/(5,2) granules with the hardware convert after pretruncate()
# (algorithm proven against the reference cast on the CPU, tests/cpp; not yet measured on the GPU -> off by default)
_PACK_HW = 0x100 if os.environ.get("CPD_B200_PACK_HW", "0") == "1" else 0
# opt-in: only CTA 0 polls the peers' flags and republishes a local go-flag (flag-round contention experiment)
_PACK_HW |= 0x200 if os.environ.get("CPD_B200_POLL_CTA0", "0") == "1" else 0
# opt-in: plain loads for the peers' wire data in the fold (do 7 of them overlap better than strong sys-scope loads?)
_PACK_HW |= 0x400 if os.environ.get("CPD_B200_WEAK_LOADS", "0") == "1" else 0
# opt-in: drop the explicit system fence in front of the flags' release stores
_PACK_HW |= 0x800 if os.environ.get("CPD_B200_NO_EXTRA_FENCE", "0") == "1" else 0
_FAULT_DELAY_RANK = int(os.environ.get("CPD_B200_FAULT_DELAY_RANK", "-1"))
_FAULT_DELAY_CYCLES = float(os.environ.get("CPD_B200_FAULT_DELAY_CYCLES", "2e8"))
ONESHOT_MAX_BYTES = 512 * 1024
MAX_TENSORS_PER_BUCKET = 1024
RAW_MAX_ELEMS = 1024


class GradReducer:
    """Low-precision gradient all-reduce of a flat gradient arena (dist_util.py:22-89 semantics, bit for bit).

    ``mode``: "fused" = one peer-memory kernel per bucket (raw / one-shot / two-shot / NVLS picked by size), "gather" =
    the reference algorithm on NCCL/Gloo (all_gather + one fold launch; fallback and oracle), "local" = world size 1.
    Buckets follow ``cfg.bucket_mb``; with ``cfg.overlap`` they launch from gradient hooks on a comm stream."""
    def __init__(self, arena: ModelArena, cfg: PrecisionConfig, group=None, mode: Optional[str] = None,
                 timeout_cycles: int = 20_000_000_000):
        self.arena = arena
        self.cfg = cfg
        self.group = group
        self.device = arena.device
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.timeout_cycles = int(timeout_cycles)
        self.is_cuda = self.device.type == "cuda"

        # wire width: narrow codes need already-quantised values (APS or prequantize)
        nat = native.wire_bits(cfg.grad_exp, cfg.grad_man)
        if cfg.wire != "auto":
            self.wire = int(cfg.wire)
            if self.wire < nat:
                raise ValueError(f"wire={self.wire} cannot hold (exp={cfg.grad_exp}, man={cfg.grad_man}); need {nat}")
        else:
            self.wire = nat
        if not (cfg.use_aps or cfg.prequantize):
            self.wire = 32          # raw fp32 gradients travel: the first cast happens inside the fold

        explicit_mode = mode is not None
        if mode is None:
            if self.world == 1:
                mode = "local"
            elif cfg.algo == "nccl" or not self.is_cuda:
                mode = "gather"
            else:
                mode = "fused"
        self.mode = mode

        ranges = arena.bucket_ranges(cfg.bucket_mb * (1 << 20))
        split = []
        for (a, b) in ranges:                      # the kernels keep per-tensor shifts in shared memory
            while b - a > MAX_TENSORS_PER_BUCKET:
                split.append((b - MAX_TENSORS_PER_BUCKET, b))
                b -= MAX_TENSORS_PER_BUCKET
            split.append((a, b))
        self.buckets: List[ChunkTable] = [arena.table(a, b) for (a, b) in split]
        self.max_T = max(b.T for b in self.buckets)
        self.max_numel = max(b.numel for b in self.buckets)
        self.shifts = [torch.zeros(b.T, dtype=torch.int32, device=self.device) for b in self.buckets]

        self.local_ws = None
        self.symm = None
        self._gather_buf = None
        if self.is_cuda:
            self.local_ws = torch.zeros(4 + 2 * MAX_TENSORS_PER_BUCKET + 32, dtype=torch.int32, device=self.device)
        if self.mode == "fused":
            from .symm import SymmetricWorkspace
            cap = max(self.max_numel * (self.wire // 8), min(self.max_numel, RAW_MAX_ELEMS) * 4)
            if cfg.algo == "raw":
                cap = self.max_numel * 4
            try:
                self.symm = SymmetricWorkspace(cap, MAX_TENSORS_PER_BUCKET, group=group)
            except RuntimeError as err:
                # no peer-memory backend (no P2P, > 8 ranks, ...): every rank fails together (the bring-up agrees on
                # each step collectively), so an auto-selected fused mode degrades to the reference algorithm on NCCL
                if explicit_mode:
                    raise
                if self.rank == 0:
                    warnings.warn(f"cpd_b200: fused all-reduce unavailable ({err}); falling back to all_gather + fold on NCCL")
                self.mode = "gather"
        if self.mode == "gather" and self.world > 1:
            self._gather_buf = torch.empty(self.world, self.max_numel, dtype=torch.float32, device=self.device)

        # overlap machinery
        self._hooks = []
        self._pending = None
        self._comm_stream = None
        self._launched = []
        if cfg.overlap and self.is_cuda and self.mode in ("fused", "gather", "local"):
            self._install_hooks()

    # ------------------------------------------------------------------ helpers --
    def _bucket_view(self, flat: torch.Tensor, b: ChunkTable) -> torch.Tensor:
        return flat[b.base:b.base + b.numel]

    def algo_for(self, b: ChunkTable) -> int:
        a = _ALGO.get(self.cfg.algo, native.ALGO_AUTO)
        if a != native.ALGO_AUTO:
            return a
        nbytes = b.numel * (self.wire // 8)
        if self.cfg.use_aps and b.numel <= RAW_MAX_ELEMS:
            return native.ALGO_ONESHOT_RAW              # latency-bound: one flag round, raw fp32 on the wire
        if self.world <= 2 or nbytes <= ONESHOT_MAX_BYTES:
            return native.ALGO_ONESHOT
        return native.ALGO_NVLS if (self.symm is not None and self.symm.has_multicast) else native.ALGO_TWOSHOT

    # ------------------------------------------------------------------- reduce --
    def reduce_bucket(self, i: int, grad_flat: Optional[torch.Tensor] = None) -> None:
        cfg, b = self.cfg, self.buckets[i]
        if _FAULT_DELAY_RANK == self.rank and self.is_cuda:
            # fault injection for tests: this rank reaches the collective late (bounded-spin paths)
            torch.cuda._sleep(int(_FAULT_DELAY_CYCLES))
        g = self._bucket_view(self.arena.grad_flat if grad_flat is None else grad_flat, b)
        e, m = cfg.grad_exp, cfg.grad_man
        if self.mode == "local":
            if cfg.is_fp32 and not cfg.use_kahan and not cfg.use_aps:
                return                                           # fp32 all-reduce over one rank: nothing to do
            native.aps_fold([g], g, b.seg_off, e, m, cfg.use_aps, cfg.use_kahan, 1.0, chunks=b.chunks,
                            workspace=self.local_ws, shifts_out=self.shifts[i], max_ctas=cfg.comm_ctas)
        elif self.mode == "fused":
            native._C.allreduce_fused_cuda(
                self.symm.buffers, g.data_ptr(), b.numel, b.chunks.data_ptr(), b.n_chunks, b.T, e, m,
                bool(cfg.use_aps), bool(cfg.use_kahan), self.algo_for(b) | _PACK_HW, bool(cfg.prequantize), self.wire,
                self.local_ws.data_ptr(), self.shifts[i].data_ptr(), int(cfg.comm_ctas), self.timeout_cycles,
                torch.cuda.current_stream(self.device).cuda_stream)
        else:  # gather: the reference algorithm, one collective + one fold launch per bucket
            if cfg.is_fp32 and not cfg.use_kahan and not cfg.use_aps:
                dist.all_reduce(g, group=self.group)             # dist_util.py:55-59
                return
            buf = self._gather_buf[:, :b.numel]
            dist.all_gather([buf[r] for r in range(self.world)], g, group=self.group)
            srcs = [buf[r] for r in range(self.world)]
            if cfg.prequantize and not cfg.use_aps:
                for s in srcs:
                    native.quantize_(s, s, e, m)
            native.aps_fold(srcs, g, b.seg_off, e, m, cfg.use_aps, cfg.use_kahan, float(self.world), chunks=b.chunks,
                            workspace=self.local_ws, shifts_out=self.shifts[i], max_ctas=cfg.comm_ctas)

    def reduce(self, grad_flat: Optional[torch.Tensor] = None) -> None:
        """Reduce every bucket now (call after backward)."""
        if self._pending is not None:
            self.finish()
            return
        for i in range(len(self.buckets)):
            self.reduce_bucket(i, grad_flat)

    # ------------------------------------------------------------------ overlap --
    def _install_hooks(self) -> None:
        self._comm_stream = torch.cuda.Stream(device=self.device, priority=-1)
        owner = {}
        for bi, b in enumerate(self.buckets):
            for t in range(b.t0, b.t1):
                owner[t] = bi
        self._bucket_need = [b.T for b in self.buckets]
        self._pending = [0] * len(self.buckets)
        self._armed = False

        def make_hook(bi):
            def hook(_param):
                if not self._armed:
                    return
                self._pending[bi] += 1
                if self._pending[bi] == self._bucket_need[bi]:
                    self._launch_async(bi)
            return hook

        for t, p in enumerate(self.arena.params):
            self._hooks.append(p.register_post_accumulate_grad_hook(make_hook(owner[t])))

    def arm(self) -> None:
        """Call before backward when overlap is on: buckets launch as soon as they fill."""
        if self._pending is not None:
            self._pending = [0] * len(self.buckets)
            self._launched = []
            self._armed = True

    def _launch_async(self, bi: int) -> None:
        cur = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._comm_stream.wait_event(ev)
        with torch.cuda.stream(self._comm_stream):
            self.reduce_bucket(bi)
        self._launched.append(bi)

    def finish(self) -> None:
        """Join the comm stream; reduce any bucket whose hooks did not all fire (e.g. unused params)."""
        if self._pending is None:
            return
        self._armed = False
        for bi in range(len(self.buckets)):
            if bi not in self._launched:
                self._launch_async(bi)
        done = torch.cuda.Event()
        done.record(self._comm_stream)
        torch.cuda.current_stream(self.device).wait_event(done)
        self._launched = []

    # ------------------------------------------------------------------- status --
    def check_status(self) -> None:
        """Raise if a fused kernel timed out waiting for a peer (host-side sync point)."""
        if self.local_ws is None:
            return
        st = int(self.local_ws[3].item()) & 0xFFFFFFFF
        if st & 0x80000000:
            phase, peer = (st >> 8) & 0xFF, st & 0xFF
            raise RuntimeError(f"cpd_b200: fused all-reduce timed out in phase {phase} waiting for rank {peer}")

    def phase_times_us(self) -> dict:
        """Per-phase durations (us) of the LAST fused launch, from the %globaltimer stamps CTA 0 leaves in the
        workspace: amax, exponent exchange, pack, data barrier, fold, gather barrier (two-shot), unpack."""
        if self.local_ws is None or self.mode != "fused":
            return {}
        t = self.local_ws[4 + 2 * MAX_TENSORS_PER_BUCKET:4 + 2 * MAX_TENSORS_PER_BUCKET + 32].cpu().view(torch.int64).tolist()
        d = lambda a, b: round((t[b] - t[a]) / 1e3, 2) if t[a] and t[b] and t[b] >= t[a] else None  # noqa: E731
        fine = {"C_first_granule": {"chunk+shift": d(4, 9), "peer_loads": d(9, 10), "decode+fold": d(10, 11), "encode+store": d(11, 12)}} \
            if t[9] >= t[4] > 0 else {}
        if t[4] and t[2] < t[0]:                    # raw variant: no exponent / pack stamps in this launch (older ones are stale)
            return {"A_amax+gridbar": d(0, 1), "push+flags": d(1, 4), "fold": d(4, 7), "total": d(0, 7),
                    "exp_exchange": None, "B_pack+gridbar": None, "data_flags": None, "C_fold": None,
                    "gather_gridbar+flags": None, "D_unpack": None}
        return {**fine, "A_amax+gridbar": d(0, 1), "exp_exchange": d(1, 2), "B_pack+gridbar": d(2, 3), "data_flags": d(3, 4), "C_fold": d(4, 5),
                "gather_gridbar+flags": d(5, 6), "D_unpack": d(6, 7) if t[6] else None, "total": d(0, 7)}

    def guard_counters(self) -> dict:
        """Cumulative counters kept by the fused kernel: tensors whose cross-rank max |g| was Inf/NaN (their shift is
        forced to 0 and the non-finite values propagate) and all-zero tensors (shift 0 instead of the reference's NaN)."""
        if self.local_ws is None or self.mode != "fused":
            return {}
        base = 4 + 2 * MAX_TENSORS_PER_BUCKET + 16
        c = self.local_ws[base:base + 2].cpu().tolist()
        return {"nonfinite_tensor_events": int(c[0]), "all_zero_tensor_events": int(c[1])}

    def remove_hooks(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
        self._pending = None
