#!/usr/bin/env python
"""Tiny driver that launches each hand-written single-GPU kernel a few times so `ncu -k regex:...` can capture it."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpd_b200.ops import native  # noqa: E402
from cpd_b200.parallel.arena import ArenaLayout  # noqa: E402
from cpd_b200.models import create_model  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
if which in ("all", "gemm_tc"):
    M = N = K = 4096
    a = torch.randn(M, K, device="cuda").bfloat16()
    bt = torch.randn(N, K, device="cuda").bfloat16()
    for _ in range(3):
        native.gemm_tc(a, bt, 5, 10, kahan=False, k_slice=256, operand_mode=native.OPERAND_BF16)
    a8, b8 = a.to(torch.float8_e4m3fn), bt.to(torch.float8_e4m3fn)
    for _ in range(3):
        native.gemm_tc(a8, b8, 4, 3, kahan=False, k_slice=512, operand_mode=native.OPERAND_E4M3)
if which in ("gemm_exact",):
    a, b = torch.randn(1024, 1024, device="cuda"), torch.randn(1024, 1024, device="cuda")
    for _ in range(2):
        native.gemm_exact(a, b, 4, 3)
if which in ("mx",):
    a, bt = torch.randn(4096, 4096, device="cuda"), torch.randn(4096, 4096, device="cuda")
    for _ in range(2):
        ac, asf = native.mx_quantize(a)
        bc, bsf = native.mx_quantize(bt)
        native.gemm_tc_mx(ac, asf, bc, bsf, 5, 10, k_slice=512)
if which in ("quantize",):
    x = torch.randn(1 << 26, device="cuda")
    o = torch.empty_like(x)
    for _ in range(3):
        native.quantize_(x, o, 4, 3)
if which in ("gemm_tc2",):
    M = N = K = 8192
    a = torch.randn(M, K, device="cuda").bfloat16()
    bt = torch.randn(N, K, device="cuda").bfloat16()
    for _ in range(3):
        native.gemm_tc(a, bt, 5, 10, kahan=False, k_slice=256, operand_mode=native.OPERAND_BF16, variant=native.GEMM_2CTA)
if which in ("all", "fold", "lars"):
    numels = [p.numel() for p in create_model("resnet18_cifar").parameters()]
    layout = ArenaLayout(numels)
    tab = layout.table("cuda")
    ws = torch.zeros(4 + 2 * 128, dtype=torch.int32, device="cuda")
    g = torch.randn(layout.total, device="cuda") * 1e-3
    out = torch.empty_like(g)
    if which in ("all", "fold"):
        for _ in range(3):
            native.aps_fold([g], out, tab.seg_off, 4, 3, True, False, 1.0, chunks=tab.chunks, workspace=ws)
    if which in ("all", "lars"):
        master, mom, model = torch.randn_like(g), torch.zeros_like(g), torch.empty_like(g)
        partials = torch.zeros(2 * tab.n_chunks, device="cuda")
        bar = torch.zeros(4, dtype=torch.int32, device="cuda")
        lr = torch.tensor([0.1], device="cuda")
        for _ in range(3):
            native.lars_step(master, model, g, mom, tab.seg_off, 0.1, 0.9, 1e-4, use_lars=True, chunks=tab.chunks, tchunk_off=tab.tchunk_off,
                             lr_dev=lr, partials=partials, barrier_ws=bar)
if which in ("stage", "conv"):
    from cpd_b200 import Quant_Conv
    x = torch.randn(64, 64, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    conv = Quant_Conv(64, 64, 3, padding=1, bias=False, exp=5, man=10, mode="tile", operands="bf16").cuda().to(memory_format=torch.channels_last)
    for _ in range(3):                      # layer-1 shape of ResNet-18/CIFAR at batch 64: staging (im2col, transposed im2col), GEMMs, split-K fold
        y = conv(x)
        y.backward(torch.ones_like(y))
if which in ("bn",):
    from cpd_b200.nn import BatchNormAct2d
    bn = BatchNormAct2d(64, relu=True).cuda()
    x = torch.randn(512, 64, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    r = torch.randn(512, 64, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    for _ in range(3):                      # the largest BatchNorm of ResNet-18/CIFAR at batch 512: 134 MB activations
        y = bn(x, r)
        y.backward(torch.ones_like(y))
if which in ("allreduce_fake",):
    # the cross-GPU kernel told world = 8 with every peer aliased to this GPU (tools/fake_world.py): ncu can replay it
    import subprocess
    sys.exit(subprocess.call([sys.executable, os.path.join(ROOT, "tools", "fake_world.py"), "--max-bytes", str(1 << 26)]))
torch.cuda.synchronize()
print("done", which)
