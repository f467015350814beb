#!/usr/bin/env python
"""Which MN-major shared-memory descriptor strides does tcgen05 want for the implicit weight gradient?  Runs
native.gemm_wgrad with candidate (lbo, sbo) byte strides against autograd's fp32 weight gradient and prints the error
of each (a wrong descriptor gives wrong numbers, not a hang)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpd_b200.ops import native  # noqa: E402

torch.manual_seed(0)
B, C, H, W, O, k = 4, 64, 16, 16, 128, 3
x = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
dy = torch.randn(B, O, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
x_nhwc = x.permute(0, 2, 3, 1).contiguous().bfloat16()
dy_rows = dy.permute(0, 2, 3, 1).reshape(-1, O).contiguous().bfloat16()
wr = torch.zeros(O, C, k, k, device="cuda", requires_grad=True)
(dwr,) = torch.autograd.grad(F.conv2d(x.bfloat16().float(), wr, None, padding=1), wr, dy.bfloat16().float())
ref = dwr.permute(0, 2, 3, 1).reshape(O, k * k * C)
for lbo, sbo in ((8192, 1024), (1024, 8192), (8192, 2048), (2048, 8192), (16, 1024), (1024, 16)):
    got = native.gemm_wgrad(dy_rows, x_nhwc, k, k, (H, W), 1, 1, 1, 8, 23, False, 0, _lbo=lbo, _sbo=sbo)
    torch.cuda.synchronize()
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    print(f"lbo={lbo:5d} sbo={sbo:5d} rel_err={err:.3e} {'<-- matches' if err < 1e-3 else ''}")
