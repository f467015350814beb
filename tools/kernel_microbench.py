#!/usr/bin/env python
"""Single-GPU micro-benchmarks of the elementwise / multi-tensor kernels with roofline fractions
(measured HBM copy bandwidth from MEASURED_PEAKS.json).  CUDA-event timed, L2 flushed between
iterations by writing a 256 MiB buffer.  Prints one JSON object."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpd_b200 import float_quantize, quant_gemm  # noqa: E402
from cpd_b200.ops import native  # noqa: E402
from cpd_b200.parallel.arena import ArenaLayout  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


_flush = None


def time_ms(fn, iters=10, warm=3, flush=True):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            _flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    hbm, tflops, src = peaks()
    res = {"peaks": {"hbm_gbs": hbm, "bf16_tflops": tflops, "source": src}, "quantize": [], "aps_fold": [], "lars": [], "gemm_exact": []}
    sizes = [1 << 24] if args.quick else [1 << 20, 1 << 24, 1 << 28]
    for n in sizes:
        x = torch.randn(n, device="cuda")
        out = torch.empty_like(x)
        for (e, m) in [(4, 3), (5, 2)] if not args.quick else [(4, 3)]:
            ms = time_ms(lambda: native.quantize_(x, out, e, m))
            gbs = 8 * n / ms / 1e6
            res["quantize"].append({"n": n, "fmt": [e, m], "ms": round(ms, 4), "GBps": round(gbs, 1), "frac_of_hbm_" + src: round(gbs / hbm, 3)})
        del x, out
    # whole-model fused kernels on the ResNet-18 layout (62 tensors, 11.17 M elements)
    from cpd_b200.models import create_model
    numels = [p.numel() for p in create_model("resnet18_cifar").parameters()]
    layout = ArenaLayout(numels)
    tab = layout.table("cuda")
    ws = torch.zeros(4 + 2 * 128, dtype=torch.int32, device="cuda")
    for W in ([1] if args.quick else [1, 8]):
        srcs = [torch.randn(layout.total, device="cuda") * 1e-3 for _ in range(W)]
        out = torch.empty(layout.total, device="cuda")
        for kahan in (False, True):
            ms = time_ms(lambda: native.aps_fold(srcs, out, tab.seg_off, 4, 3, True, kahan, float(W), chunks=tab.chunks, workspace=ws))
            byts = layout.total * 4 * (W + 1)          # algorithmic: read W arenas once (2nd pass hits L2), write one
            res["aps_fold"].append({"W": W, "kahan": kahan, "elems": layout.total, "ms": round(ms, 4), "GBps_algorithmic": round(byts / ms / 1e6, 1),
                                    "frac_of_hbm_" + src: round(byts / ms / 1e6 / hbm, 3)})
    master, grad, mom, model = (torch.randn(layout.total, device="cuda") for _ in range(4))
    partials = torch.zeros(2 * tab.n_chunks, device="cuda")
    bar = torch.zeros(4, dtype=torch.int32, device="cuda")
    lr_dev = torch.tensor([0.1], device="cuda")
    for use_lars in (True, False):
        ms = time_ms(lambda: native.lars_step(master, model, grad, mom, tab.seg_off, 0.1, 0.9, 1e-4, use_lars=use_lars, chunks=tab.chunks,
                                              tchunk_off=tab.tchunk_off, lr_dev=lr_dev, partials=partials, barrier_ws=bar))
        byts = layout.total * 4 * 6                     # read w,g,m ; write w,m,model
        res["lars"].append({"use_lars": use_lars, "ms": round(ms, 4), "GBps_algorithmic": round(byts / ms / 1e6, 1),
                            "frac_of_hbm_" + src: round(byts / ms / 1e6 / hbm, 3)})
    if not args.quick:
        for n in (64, 1024):
            a, b = torch.randn(n, n, device="cuda"), torch.randn(n, n, device="cuda")
            ms = time_ms(lambda: quant_gemm(a, b, man=3, exp=4, mode="exact"), iters=5, flush=False)
            res["gemm_exact"].append({"n": n, "ms": round(ms, 4), "GFLOPs": round(2 * n ** 3 / ms / 1e6, 1)})
    # tensor-core tile mode end to end: fp32 operands in, fused gather+cast staging (one launch per operand) + GEMM,
    # against the same GEMM on pre-staged operands (kernel only) and cuBLAS bf16 on the same box
    res["gemm_tile_e2e"], res["operand_stage"] = [], []
    for n in ([4096] if args.quick else [4096, 8192]):
        a, b = torch.randn(n, n, device="cuda"), torch.randn(n, n, device="cuda")
        ao, bo = native.stage_matrix(a, native.OPERAND_BF16), native.stage_matrix(b.t(), native.OPERAND_BF16)
        out = torch.empty(n, n, device="cuda")
        ms_k = time_ms(lambda: native.gemm_tc(ao, bo, 5, 10, False, 0, native.OPERAND_BF16, None, out=out, k=n), iters=6, flush=False)
        ms_e = time_ms(lambda: quant_gemm(a, b, man=10, exp=5, mode="tile", operands="bf16", k_slice=0), iters=6, flush=False)
        ab, bb = a.bfloat16(), b.bfloat16()
        ms_c = time_ms(lambda: torch.matmul(ab, bb), iters=6, flush=False)
        ms_s = time_ms(lambda: native.stage_matrix(a, native.OPERAND_BF16), iters=6)
        ms_st = time_ms(lambda: native.stage_matrix(b.t(), native.OPERAND_BF16), iters=6)
        fl = 2 * n ** 3
        res["gemm_tile_e2e"].append({"n": n, "kernel_only_ms": round(ms_k, 4), "kernel_only_TFLOPs": round(fl / ms_k / 1e9, 1),
                                     "e2e_fp32_in_ms": round(ms_e, 4), "e2e_TFLOPs": round(fl / ms_e / 1e9, 1),
                                     "e2e_over_kernel_only": round(ms_k / ms_e, 3), "cublas_bf16_ms": round(ms_c, 4),
                                     "cublas_bf16_TFLOPs": round(fl / ms_c / 1e9, 1)})
        res["operand_stage"].append({"n": n, "row_major_ms": round(ms_s, 4), "row_major_GBps": round(6 * n * n / ms_s / 1e6, 1),
                                     "transposed_ms": round(ms_st, 4), "transposed_GBps": round(6 * n * n / ms_st / 1e6, 1),
                                     "frac_of_hbm_" + src: round(6 * n * n / ms_s / 1e6 / hbm, 3)})
        del a, b, ao, bo, out, ab, bb
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
