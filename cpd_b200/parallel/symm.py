"""Symmetric (peer-mapped) workspace for the fused all-reduce.

One region per process group, carved into wire buffers, gather buffers, the APS
exponent mailbox and the signal pad.  Three bring-up backends, tried in order:
  ``vmm``   own allocator (csrc/comm/symm_mem.cpp): cuMemCreate + POSIX-fd exchange over
            abstract unix sockets + cuMulticast (NVLS) view
  ``torch`` torch.distributed._symmetric_memory (buffer_ptrs / multicast_ptr)
  ``ipc``   cudaMalloc + legacy CUDA IPC handles (no multicast)
The kernels only ever see raw pointers, so they are agnostic to the backend.
"""
from __future__ import annotations

import os
import uuid
from typing import List, Optional

import torch
import torch.distributed as dist

from .._ext import load

_C = load()

FLAG_WORDS = 3 * 16
_ALIGN = 4096


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


class SymmetricWorkspace:
    def __init__(self, cap_bytes: int, max_tensors: int = 1024, group=None, backend: str = "auto",
                 want_multicast: bool = True):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("SymmetricWorkspace needs an initialised process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 8:
            raise RuntimeError("fused all-reduce supports up to 8 GPUs per NVSwitch domain")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.cap_bytes = _round_up(int(cap_bytes), _ALIGN)
        self.max_tensors = int(max_tensors)
        self.off_data = 0
        self.off_gather = self.off_data + 2 * self.cap_bytes
        self.off_mailbox = self.off_gather + 2 * self.cap_bytes
        self.off_flags = self.off_mailbox + _round_up(2 * 8 * self.max_tensors * 4, _ALIGN)      # [parity][source rank][tensor]
        self.total = self.off_flags + _round_up(FLAG_WORDS * 4, _ALIGN)
        self.backend = None
        self.log: List[str] = []
        self._keep = []
        ptrs, mc = None, 0
        order = [backend] if backend != "auto" else os.environ.get("CPD_B200_SYMM", "vmm,torch,ipc").split(",")
        last_err = None
        for b in order:
            try:
                ptrs, mc = getattr(self, "_open_" + b)(want_multicast)
                ok = torch.tensor([1], device=self.device)
            except Exception as e:  # noqa: BLE001 - report and try the next backend
                last_err = e
                self.log.append(f"{b}: {type(e).__name__}: {e}")
                ok = torch.tensor([0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 1:
                self.backend = b
                break
            ptrs = None
        if ptrs is None:
            raise RuntimeError("cpd_b200: no symmetric-memory backend available: " + "; ".join(self.log)) from last_err
        # multicast only if every rank got it
        has_mc = torch.tensor([1 if mc else 0], device=self.device)
        dist.all_reduce(has_mc, op=dist.ReduceOp.MIN, group=group)
        self.mc_base = mc if int(has_mc.item()) == 1 else 0
        self.ptrs = ptrs
        self.buffers = _C.CommBuffers(
            self.rank, self.world,
            [p + self.off_data for p in ptrs], [p + self.off_gather for p in ptrs],
            [p + self.off_mailbox for p in ptrs], [p + self.off_flags for p in ptrs],
            (self.mc_base + self.off_gather) if self.mc_base else 0, self.cap_bytes, self.max_tensors,
            (self.mc_base + self.off_data) if self.mc_base else 0)
        # device-local protocol state of THIS workspace (grid barrier, launch epoch, status word, amax[2][T], stamps):
        # the epoch must live with the flags it is compared against, so every reducer that drives this workspace
        # shares it
        self.local_ws = torch.zeros(4 + 2 * self.max_tensors + 32, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize()
        dist.barrier(group=group)

    @property
    def has_multicast(self) -> bool:
        return bool(self.mc_base)

    def close(self) -> None:
        """Collective teardown: every rank of the group must call it, after its last kernel on the region.  Device sync
        + barrier (no rank may still read a peer mapping), release (vmm: unmap / multicast unbind / handle release;
        ipc: close peer handles, free), barrier.  Idempotent."""
        if self.ptrs is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for item in self._keep:
            if isinstance(item, tuple) and item[0] == "ipc":
                for r, p in enumerate(self.ptrs):
                    if r != self.rank:
                        _C.ipc_close(p)
                _C.ipc_free(item[1])
            elif hasattr(item, "close") and hasattr(item, "bind_multicast"):
                item.close()
                if item.log():
                    self.log.append("close: " + item.log())
        self._keep = []
        self.ptrs, self.buffers, self.mc_base = None, None, 0
        dist.barrier(group=self.group)

    # -- backends -------------------------------------------------------------------
    def _agree(self, step, what: str):
        """Run ``step`` on every rank; raise on ALL ranks if it failed on any (keeps the
        collectives of the bring-up aligned even when one rank hits an error)."""
        err = None
        try:
            out = step()
        except Exception as e:  # noqa: BLE001
            err, out = e, None
        ok = torch.tensor([0 if err else 1], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) != 1:
            raise RuntimeError(f"{what} failed on at least one rank" + (f": {err}" if err else ""))
        return out

    def _open_vmm(self, want_mc: bool):
        sess = [uuid.uuid4().hex[:12] if self.rank == 0 else None]
        dist.broadcast_object_list(sess, src=dist.get_global_rank(self.group, 0) if self.group else 0, group=self.group)
        reg = self._agree(lambda: _C.SymmRegion(self.total, self.rank, self.world, self.device.index, sess[0], want_mc),
                          "SymmRegion()")
        self._agree(reg.open, "SymmRegion.open")
        self._agree(reg.exchange, "SymmRegion.exchange")
        self._agree(reg.bind_multicast, "SymmRegion.bind_multicast")
        if reg.log():
            self.log.append("vmm: " + reg.log())
        self._keep.append(reg)
        return [reg.ptr(r) for r in range(self.world)], reg.mc_ptr()

    def _open_torch(self, want_mc: bool):
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(self.total, dtype=torch.uint8, device=self.device)
        grp = self.group if self.group is not None else dist.group.WORLD
        hdl = symm_mem.rendezvous(t, group=grp.group_name)
        t.zero_()
        torch.cuda.synchronize()
        self._keep += [t, hdl]
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0) if want_mc else 0
        return [int(p) for p in hdl.buffer_ptrs], mc

    def _open_ipc(self, want_mc: bool):
        base = _C.ipc_alloc(self.total)
        handle = _C.ipc_export(base)
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, handle, group=self.group)
        ptrs = []
        for r in range(self.world):
            ptrs.append(base if r == self.rank else _C.ipc_open(handles[r]))
        self._keep.append(("ipc", base))
        return ptrs, 0
