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
         "nvls": native.ALGO_NVLS, "raw": native.ALGO_ONESHOT_RAW, "nvls_sum": native.ALGO_NVLS_SUM,
         "nvls_fast": native.ALGO_NVLS_FAST}

# Protocol switches of the fused kernel (upper bits of the algo word).  The hardware pack (bit 8: scale + reference cast
# + pack of a granule through pretruncate() + F2FP.SATFINITE, proven equal to the software cast on all 2^32 patterns
# below the overflow band, tools/exhaustive_cast_check.sh) is ON by default; CPD_B200_PACK_HW=0 switches it off.
_PROTO = 0x100 if os.environ.get("CPD_B200_PACK_HW", "1") == "1" else 0
# No explicit __threadfence_system() in front of a flag's st.release.sys (bit 11): the release is cumulative over what
# CTA 0 observed through the grid barrier, so the fence is redundant; measured -1.5..-2 us per flag round (fake world W = 8:
# one-shot 1 MB 27.5 -> 23.6 us; 2 real GPUs 24.0 -> 22.2 us), 91 cases + specials bit-exact on 2 GPUs.  ON by default;
# CPD_B200_NO_EXTRA_FENCE=0 restores the belt-and-braces fence.
_PROTO |= 0x800 if os.environ.get("CPD_B200_NO_EXTRA_FENCE", "1") == "1" else 0
# experiments that did not pay: only CTA 0 polls the peers' flags (+0.5 us), plain instead of system-scope loads of the
# peers' wire data (no gain on real peers)
_PROTO |= 0x200 if os.environ.get("CPD_B200_POLL_CTA0", "0") == "1" else 0
_PROTO |= 0x400 if os.environ.get("CPD_B200_WEAK_LOADS", "0") == "1" else 0
_FAULT_DELAY_RANK = int(os.environ.get("CPD_B200_FAULT_DELAY_RANK", "-1"))
_FAULT_DELAY_CYCLES = float(os.environ.get("CPD_B200_FAULT_DELAY_CYCLES", "2e8"))
MAX_TENSORS_PER_BUCKET = 1024
NVLS_SUM_MAX_BYTES = 128 << 20
# a peer that does not show up within this many SM cycles (~5 min at 1.9 GHz; NCCL's watchdog default is 10 min) makes
# the kernel poison its outputs (NaN) and raise the status word; tests shorten it
DEFAULT_TIMEOUT_CYCLES = 600_000_000_000


class Thresholds:
    """Which fused variant serves which bucket size.  Derived from the discovered topology (per-GPU NVLink bandwidth,
    multicast) with the latency constants measured on 8 x B200 (profiles/allreduce_*), or overridden by a calibration
    run on the live process group (``parallel.calibrate``).

    raw (ONE flag round, fp32 on the wire) wins while its extra wire bytes cost less than the second flag round of the
    packed variants: (W - 1) * 4 B * n / link_bw  <=  t_flag_round.  The packed one-shot ((W - 1) B per element in)
    hands over to two-shot / NVLS (2 (W - 1) / W B per element) when the byte difference outweighs the third flag
    round + grid barrier of the gather half."""

    FLAG_ROUND_US = 6.0       # one flag round + grid barrier on real peers (phase stamps: 4-7 us at W = 2, 7-9 at W = 8)
    LINK_DERATE = 0.5         # share of the nominal link rate a latency-exposed peer read / fold sustains

    def __init__(self, world: int, link_gbs: float = 900.0, multicast: bool = True, raw_max_elems: Optional[int] = None,
                 oneshot_max_bytes: Optional[int] = None, source: str = "topology"):
        self.world, self.link_gbs, self.multicast, self.source = world, float(link_gbs) or 900.0, multicast, source
        per_us = self.link_gbs * 1e3 * self.LINK_DERATE   # bytes per microsecond per direction, derated
        w = max(world, 2)
        if raw_max_elems is None:
            # W = 2: ~675 K elements (measured crossover between 256 K and 1 M); W = 8: ~96 K
            raw_max_elems = int(self.FLAG_ROUND_US * per_us / (4.0 * (w - 1)))
            raw_max_elems = max(1024, min(raw_max_elems, 1 << 20))
        if oneshot_max_bytes is None:
            # (W-1) - 2(W-1)/W bytes per element saved by two-shot vs ~1.5 flag rounds it adds
            saved = (w - 1) - 2.0 * (w - 1) / w
            oneshot_max_bytes = (1 << 40) if saved <= 0 else int(1.5 * self.FLAG_ROUND_US * per_us / saved)
            oneshot_max_bytes = max(64 * 1024, min(oneshot_max_bytes, 8 << 20))
        self.raw_max_elems, self.oneshot_max_bytes = int(raw_max_elems), int(oneshot_max_bytes)

    def to_dict(self) -> dict:
        return dict(world=self.world, link_gbs=self.link_gbs, multicast=self.multicast, raw_max_elems=self.raw_max_elems,
                    oneshot_max_bytes=self.oneshot_max_bytes, source=self.source)


_CALIBRATED = {}      # world size -> Thresholds installed by parallel.calibrate.calibrate()


def default_thresholds(world: int) -> Thresholds:
    if world in _CALIBRATED:
        return _CALIBRATED[world]
    link, mc = 900.0, True
    try:
        from .dist import discover_topology
        topo = discover_topology()
        link = topo.est_link_gbs or 900.0
        mc = bool(topo.multicast)
    except Exception:  # noqa: BLE001 - CPU-only box: constants are unused there
        pass
    return Thresholds(world, link, mc)


class GradReducer:
    """Low-precision gradient all-reduce of a flat gradient arena (dist_util.py:22-89 semantics, bit for bit).

    ``mode``: "fused" = one peer-memory kernel per bucket (raw / one-shot / two-shot / NVLS picked by size), "gather" =
    the reference algorithm on NCCL/Gloo (all_gather + one fold launch; fallback and oracle), "local" = world size 1.
    Buckets follow ``cfg.bucket_mb``; with ``cfg.overlap`` they launch from gradient hooks on a comm stream."""
    def __init__(self, arena: ModelArena, cfg: PrecisionConfig, group=None, mode: Optional[str] = None,
                 timeout_cycles: int = DEFAULT_TIMEOUT_CYCLES, thresholds: Optional[Thresholds] = None, symm=None):
        self.arena = arena
        self.cfg = cfg
        self.group = group
        self.device = arena.device
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.timeout_cycles = int(timeout_cycles)
        self.is_cuda = self.device.type == "cuda"
        self.thresholds = thresholds or default_thresholds(self.world)
        self.rounding = native.rounding_code(cfg.rounding)
        # (8,23) without Kahan / APS is a plain sum: reference = NCCL all_reduce (dist_util.py:55-59)
        self.plain_fp32 = cfg.is_fp32 and not cfg.use_kahan and not cfg.use_aps

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

        # overlap needs several buckets: default 8 MB of fp32 gradients each (a bucket's reduction then hides behind
        # the backward of the layers below it); without overlap one launch covers the whole model
        bucket_mb = cfg.bucket_mb if cfg.bucket_mb > 0 else (8.0 if cfg.overlap else 0.0)
        ranges = arena.bucket_ranges(bucket_mb * (1 << 20))
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
        self._owns_symm = True
        self._gather_buf = None
        if self.is_cuda:
            self.local_ws = torch.zeros(4 + 2 * MAX_TENSORS_PER_BUCKET + 32, dtype=torch.int32, device=self.device)
        if self.mode == "fused":
            from .symm import SymmetricWorkspace
            cap = max(self.max_numel * (self.wire // 8), min(self.max_numel, self.thresholds.raw_max_elems) * 4)
            if cfg.algo == "raw":
                cap = self.max_numel * 4
            try:
                if symm is not None:                       # a caller-owned workspace (calibration, tools): must be big enough
                    if symm.cap_bytes < cap:
                        raise ValueError(f"shared symmetric workspace too small: {symm.cap_bytes} < {cap} bytes")
                    self.symm, self._owns_symm = symm, False
                else:
                    self.symm = SymmetricWorkspace(cap, MAX_TENSORS_PER_BUCKET, group=group)
                self.local_ws = self.symm.local_ws          # epoch / status / barrier words belong to the workspace
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

        self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory() if self.is_cuda else None
        # set by TrainEngine(steal_grads=True): an overlapped bucket first moves its parameters' gradients into the arena
        self.gather_before_launch = False
        # overlap machinery
        self._hooks = []
        self._pending = None
        self._last_ready_bucket = -1
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
        th = self.thresholds
        has_mc = self.symm is not None and self.symm.has_multicast
        if self.plain_fp32:
            # fp32 on the wire: the one-shot costs what the raw variant costs, so it keeps the same size range; above it
            # the NVSwitch does the sum (multimem.ld_reduce) where there is one, a two-shot fp32 fold otherwise
            if b.numel <= th.raw_max_elems:
                return native.ALGO_ONESHOT
            if has_mc and self.world > 2:
                return native.ALGO_NVLS_SUM
            return native.ALGO_ONESHOT if self.world <= 2 else native.ALGO_TWOSHOT
        nbytes = b.numel * (self.wire // 8)
        if self.cfg.use_aps and b.numel <= th.raw_max_elems:
            return native.ALGO_ONESHOT_RAW              # latency-bound: one flag round, raw fp32 on the wire
        if self.world <= 2 or nbytes <= th.oneshot_max_bytes:
            return native.ALGO_ONESHOT
        return native.ALGO_NVLS if has_mc else native.ALGO_TWOSHOT

    def ctas_for(self, i: int) -> int:
        """Grid of bucket ``i``'s launch.  Overlapped buckets co-run with backward kernels: a SMALL grid (cfg.comm_ctas,
        default 32) leaves the SMs to them and stays co-resident (the kernel's grid barrier needs that); the bucket
        that becomes ready LAST has nothing left to overlap with and takes the full grid."""
        if self._pending is None:
            return int(self.cfg.comm_ctas)
        if i == self._last_ready_bucket:
            return 0
        return int(self.cfg.comm_ctas) or 32

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
                            workspace=self.local_ws, shifts_out=self.shifts[i], max_ctas=self.ctas_for(i),
                            rounding=self.rounding)
        elif self.mode == "fused":
            if self.plain_fp32 and b.numel * 4 > NVLS_SUM_MAX_BYTES and cfg.algo == "auto":
                # plain fp32 sum of a very large bucket: the in-switch variant's copy-in / copy-out passes lose to NCCL's
                # in-place NVLS all_reduce above ~128 MB (8 x B200: 256 MB 895 vs 717 us) - which IS the reference path here
                dist.all_reduce(g, group=self.group)
                return
            native._C.allreduce_fused_cuda(
                self.symm.buffers, g.data_ptr(), b.numel, b.chunks.data_ptr(), b.n_chunks, b.T, e, m, self.rounding,
                bool(cfg.use_aps), bool(cfg.use_kahan), self.algo_for(b) | _PROTO, bool(cfg.prequantize), self.wire,
                self.local_ws.data_ptr(), self.shifts[i].data_ptr(), self.ctas_for(i), self.timeout_cycles,
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
                    native.quantize_(s, s, e, m, cfg.rounding)
            native.aps_fold(srcs, g, b.seg_off, e, m, cfg.use_aps, cfg.use_kahan, float(self.world), chunks=b.chunks,
                            workspace=self.local_ws, shifts_out=self.shifts[i], max_ctas=self.ctas_for(i),
                            rounding=self.rounding)

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
        # gradients become ready from the last parameter to the first: the bucket holding tensor 0 fills last
        self._last_ready_bucket = owner[0]
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
        if self.gather_before_launch:
            b = self.buckets[bi]
            self.arena.gather_grads(b.t0, b.t1)            # on the compute stream, behind the kernels that wrote them
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
        """Raise if a fused kernel timed out waiting for a peer (host-side sync point).  The kernel has then poisoned
        the gradients it produced (NaN), so the step cannot have applied stale peer data silently."""
        if self.local_ws is None:
            return
        self._raise_on_status(int(self.local_ws[3].item()))

    def status_async(self) -> Optional[torch.Tensor]:
        """Start a 4-byte device->pinned-host copy of the status word on the current stream (the training engine issues
        it next to the loss read-back of every step and calls ``check_status_host`` after its event sync)."""
        if self.local_ws is None or self.mode != "fused":
            return None
        self._status_host.copy_(self.local_ws[3:4], non_blocking=True)
        return self._status_host

    def check_status_host(self) -> None:
        if self._status_host is not None and self.mode == "fused":
            self._raise_on_status(int(self._status_host[0]))

    def _raise_on_status(self, st: int) -> None:
        st &= 0xFFFFFFFF
        if st & 0x80000000:
            phase, peer = (st >> 8) & 0xFF, st & 0xFF
            raise RuntimeError(f"cpd_b200: fused all-reduce timed out in phase {phase} waiting for rank {peer} "
                               f"(rank {self.rank} poisoned its gradients with NaN)")

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

    def close(self) -> None:
        """Release the symmetric workspace (collective: every rank calls it).  The reducer is unusable afterwards."""
        self.remove_hooks()
        if self.symm is not None:
            if self._owns_symm:
                self.symm.close()
            self.symm = None
            self.mode = "closed"

    def remove_hooks(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
        self._pending = None
