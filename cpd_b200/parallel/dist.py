"""Process-group bring-up and topology discovery.

Reference: CPDtorch/utils/dist_util.py:96-131 — SLURM-only (site-specific host-name
parsing, fixed port 12345), NCCL only, no topology discovery.  Here ``dist_init``
understands torchrun (RANK/LOCAL_RANK/WORLD_SIZE), SLURM, OpenMPI and the
single-process case, picks NCCL on GPUs and Gloo on CPU, binds the process to its
local GPU and records the box's GPU / NVLink / NVSwitch topology (native discovery in
csrc/comm/topo.cpp) which drives the all-reduce algorithm choice.
"""
from __future__ import annotations

import datetime
import os
import re
import socket
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from ..ops import native


@dataclass
class Topology:
    device_count: int = 0
    devices: list = field(default_factory=list)
    p2p: list = field(default_factory=list)
    nvlink: list = field(default_factory=list)
    multicast: bool = False
    raw: dict = field(default_factory=dict)

    @property
    def all_p2p(self) -> bool:
        n = self.device_count
        return n > 1 and all(self.p2p[i][j]["access"] for i in range(n) for j in range(n) if i != j)

    @property
    def nvlink_links(self) -> int:
        return min((d.get("active_links", 0) for d in self.nvlink), default=0)

    @property
    def est_link_gbs(self) -> float:
        """Per-direction NVLink bandwidth estimate per GPU (GB/s): links x per-link rate.  NVML reports the link
        generation as an enum (nvmlNvlinkVersion_t: 1 = 1.0, 2 = 2.0, 3 = 2.2, 4 = 3.0, 5 = 3.1, 6 = 4.0, 7 = 5.0);
        NVLink 5 (B200) moves 50 GB/s per link per direction, generations 2-4 move 25, NVLink 1 moves 20."""
        ver = max((d.get("version", 0) for d in self.nvlink), default=0)
        per_link = 50.0 if ver >= 7 else (20.0 if ver == 1 else 25.0)
        return self.nvlink_links * per_link

    def summary(self) -> str:
        if not self.device_count:
            return "no CUDA devices"
        name = self.devices[0]["name"] if self.devices else "?"
        return (f"{self.device_count}x {name}, p2p={'full' if self.all_p2p else 'partial'}, "
                f"nvlink links/GPU={self.nvlink_links} (~{self.est_link_gbs:.0f} GB/s/dir), multicast={self.multicast}")


_TOPOLOGY: Optional[Topology] = None
_STATE = {"rank": 0, "world_size": 1, "local_rank": 0, "launcher": "single", "initialized_here": False}


def discover_topology(refresh: bool = False) -> Topology:
    global _TOPOLOGY
    if _TOPOLOGY is None or refresh:
        raw = native.topology()
        _TOPOLOGY = Topology(device_count=raw.get("device_count", 0), devices=raw.get("devices", []),
                             p2p=raw.get("p2p", []), nvlink=raw.get("nvlink", []),
                             multicast=bool(raw.get("multicast", False)), raw=raw)
    return _TOPOLOGY


def _first_slurm_host(nodelist: str) -> str:
    """First host of a SLURM nodelist such as ``gpu[03-05,07]`` or ``nodeA,nodeB``."""
    m = re.match(r"^([^\[,]+)(?:\[([^\]]+)\])?", nodelist)
    if not m:
        return nodelist
    prefix, inner = m.group(1), m.group(2)
    if inner is None:
        return prefix
    first = re.split(r"[,-]", inner)[0]
    return prefix + first


def detect_launcher(env=None) -> dict:
    """Rank / world / master from the environment: torchrun > SLURM > OpenMPI > single."""
    env = os.environ if env is None else env
    if "RANK" in env and "WORLD_SIZE" in env:
        return dict(launcher="torchrun", rank=int(env["RANK"]), world_size=int(env["WORLD_SIZE"]),
                    local_rank=int(env.get("LOCAL_RANK", env["RANK"])),
                    master_addr=env.get("MASTER_ADDR", "127.0.0.1"), master_port=int(env.get("MASTER_PORT", 29500)))
    if "SLURM_PROCID" in env and "SLURM_NTASKS" in env:
        host = _first_slurm_host(env.get("SLURM_NODELIST", env.get("SLURM_JOB_NODELIST", "127.0.0.1")))
        return dict(launcher="slurm", rank=int(env["SLURM_PROCID"]), world_size=int(env["SLURM_NTASKS"]),
                    local_rank=int(env.get("SLURM_LOCALID", 0)), master_addr=env.get("MASTER_ADDR", host),
                    master_port=int(env.get("MASTER_PORT", 12345)))
    if "OMPI_COMM_WORLD_RANK" in env:
        return dict(launcher="openmpi", rank=int(env["OMPI_COMM_WORLD_RANK"]), world_size=int(env["OMPI_COMM_WORLD_SIZE"]),
                    local_rank=int(env.get("OMPI_COMM_WORLD_LOCAL_RANK", 0)),
                    master_addr=env.get("MASTER_ADDR", "127.0.0.1"), master_port=int(env.get("MASTER_PORT", 12345)))
    return dict(launcher="single", rank=0, world_size=1, local_rank=0, master_addr="127.0.0.1",
                master_port=int(env.get("MASTER_PORT", 0)) or _free_port())


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def dist_init(backend: Optional[str] = None, timeout_s: int = 600, verbose: bool = False,
              calibrate: Optional[bool] = None) -> Tuple[int, int]:
    """Initialise ``torch.distributed`` and return ``(rank, world_size)`` like the reference.  On GPUs the box's
    topology is discovered (it sets the all-reduce variant thresholds); ``calibrate=True`` (or env
    CPD_B200_CALIBRATE=1) additionally measures the variants on the live group once and installs the measured
    thresholds (parallel/calibrate.py, ~1 s)."""
    if dist.is_available() and dist.is_initialized():
        _STATE.update(rank=dist.get_rank(), world_size=dist.get_world_size())
        return _STATE["rank"], _STATE["world_size"]
    info = detect_launcher()
    use_cuda = torch.cuda.is_available()
    if backend is None:
        backend = "nccl" if use_cuda else "gloo"
    os.environ.setdefault("MASTER_ADDR", str(info["master_addr"]))
    os.environ.setdefault("MASTER_PORT", str(info["master_port"]))
    os.environ["RANK"] = str(info["rank"])
    os.environ["WORLD_SIZE"] = str(info["world_size"])
    kwargs = {}
    if use_cuda:
        ndev = torch.cuda.device_count()
        local = info["local_rank"] % max(ndev, 1)
        torch.cuda.set_device(local)
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
    dist.init_process_group(backend=backend, rank=info["rank"], world_size=info["world_size"],
                            timeout=datetime.timedelta(seconds=timeout_s), **kwargs)
    _STATE.update(rank=info["rank"], world_size=info["world_size"], local_rank=info["local_rank"],
                  launcher=info["launcher"], initialized_here=True)
    if use_cuda:
        topo = discover_topology()
        if verbose and info["rank"] == 0:
            print(f"[cpd_b200] launcher={info['launcher']} world={info['world_size']} backend={backend}; {topo.summary()}")
        if calibrate is None:
            calibrate = os.environ.get("CPD_B200_CALIBRATE", "0") == "1"
        if calibrate and info["world_size"] > 1 and backend == "nccl":
            from .calibrate import calibrate as _calibrate
            th = _calibrate(verbose=verbose)
            if verbose and info["rank"] == 0:
                print(f"[cpd_b200] calibrated all-reduce thresholds: {th.to_dict()}")
    return info["rank"], info["world_size"]


def get_rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shutdown() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def simple_group_split(world_size: int, rank: int, num_groups: int):
    """Split the world into ``num_groups`` contiguous process groups and return the one containing
    ``rank`` (reference: train_util.py:11-18, a helper the reference defines but never calls)."""
    size = world_size // num_groups
    groups = [dist.new_group(list(range(i * size, (i + 1) * size))) for i in range(num_groups)]
    return groups[rank // size]
