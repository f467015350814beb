"""One precision / reduction configuration object shared by every trainer and API.

The reference passes precision around as two loose ints and spells its flags
differently in every script (``--use_APS`` vs ``--use-APS``, ``--emulate_node`` vs
``--emulate-node``).  ``PrecisionConfig`` is the single source of truth
and ``add_precision_args`` accepts both spellings.
"""
from __future__ import annotations

import argparse
from dataclasses import dataclass, asdict


@dataclass
class PrecisionConfig:
    grad_exp: int = 5
    grad_man: int = 2
    use_aps: bool = False
    use_kahan: bool = False
    emulate_node: int = 1
    rounding: str = "reference"      # "reference" (bit-exact with CPD) | "ieee"
    algo: str = "auto"               # auto | oneshot | twoshot | nvls | raw | nvls_sum (fp32, in-switch) | nvls_fast (fp8 in-switch, NOT bit-exact) | nccl
    wire: str = "auto"               # auto | 8 | 16 | 32   (bits per element on NVLink)
    prequantize: bool = False        # without APS: cast before packing (enables a narrow wire; not reference-exact)
    bucket_mb: float = 0.0           # 0 = one bucket (whole model in one launch); with overlap: 8 MB buckets
    overlap: bool = False            # launch buckets from autograd hooks on a comm stream, small grids (comm_ctas)
    comm_ctas: int = 0               # CTAs of a reduction launch: 0 = full grid (overlapped buckets: 32)

    def __post_init__(self):
        if not (1 <= self.grad_exp <= 8 and 0 <= self.grad_man <= 23):
            raise ValueError(f"unsupported format (exp={self.grad_exp}, man={self.grad_man}); exp<=8, man<=23")
        if self.emulate_node < 1:
            raise ValueError("emulate_node must be >= 1")
        if self.algo not in ("auto", "oneshot", "twoshot", "nvls", "raw", "nvls_sum", "nvls_fast", "nccl"):
            raise ValueError(f"unknown algo {self.algo!r}")
        self.wire = str(self.wire)
        if self.wire not in ("auto", "8", "16", "32"):
            raise ValueError(f"unknown wire {self.wire!r}")

    @property
    def is_fp32(self) -> bool:
        return self.grad_exp == 8 and self.grad_man == 23

    def to_dict(self) -> dict:
        return asdict(self)

    @classmethod
    def from_args(cls, args) -> "PrecisionConfig":
        return cls(grad_exp=args.grad_exp, grad_man=args.grad_man, use_aps=bool(args.use_APS),
                   use_kahan=bool(getattr(args, "use_kahan", False)), emulate_node=int(args.emulate_node),
                   rounding=getattr(args, "rounding", "reference"), algo=getattr(args, "algo", "auto"),
                   wire=str(getattr(args, "wire", "auto")), bucket_mb=float(getattr(args, "bucket_mb", 0.0)),
                   overlap=bool(getattr(args, "overlap", False)), comm_ctas=int(getattr(args, "comm_ctas", 0)))


def add_precision_args(p: argparse.ArgumentParser, grad_exp: int = 5, grad_man: int = 2) -> None:
    """Legacy-compatible flags (mix.py:29-42, dawn.py:11-24, main.py:21-53), both spellings."""
    p.add_argument("--grad_exp", "--grad-exp", dest="grad_exp", type=int, default=grad_exp)
    p.add_argument("--grad_man", "--grad-man", dest="grad_man", type=int, default=grad_man)
    p.add_argument("--use_APS", "--use-APS", "--use_aps", "--use-aps", dest="use_APS", action="store_true")
    p.add_argument("--use_kahan", "--use-kahan", dest="use_kahan", action="store_true")
    p.add_argument("--use_lars", "--use-lars", dest="use_lars", action="store_true")
    p.add_argument("--emulate_node", "--emulate-node", dest="emulate_node", type=int, default=1)
    p.add_argument("--rounding", choices=["reference", "ieee"], default="reference")
    p.add_argument("--algo", choices=["auto", "oneshot", "twoshot", "nvls", "raw", "nvls_sum", "nvls_fast", "nccl"], default="auto")
    p.add_argument("--wire", choices=["auto", "8", "16", "32"], default="auto")
    p.add_argument("--bucket_mb", "--bucket-mb", dest="bucket_mb", type=float, default=0.0)
    p.add_argument("--overlap", action="store_true")
    p.add_argument("--comm_ctas", "--comm-ctas", dest="comm_ctas", type=int, default=0)
    p.add_argument("--profile_trace", "--profile-trace", dest="profile_trace", default="",
                   help="write a torch.profiler chrome trace of a few steady-state steps to this path (rank 0)")
