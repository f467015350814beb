#!/usr/bin/env python
"""GEMM + quantize micro-benchmarks vs the reference's own kernels and cuBLAS (1 GPU).
Fills the BASELINE.md rows `float_quantize GB/s` and `quant_gemm GFLOP/s`.  CUDA-event timed,
median of N, L2 flushed for the bandwidth-bound kernels.  Prints one JSON object."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from kernel_microbench import time_ms, peaks  # noqa: E402
from cpd_b200 import float_quantize, quant_gemm  # noqa: E402
from cpd_b200.ops import native  # noqa: E402


def load_reference():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "CPDtorch")):
        return None
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    from prebuild_ref_ext import ref_env
    ref_env()
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_qf", os.path.join(ref, "CPDtorch", "quant", "quant_function.py"))
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
        return mod
    except Exception as e:  # noqa: BLE001
        print("reference extension unavailable:", e, file=sys.stderr)
        return None


def main():
    quick = "--quick" in sys.argv          # tensor-core GEMM only: three shapes, bf16 operands, three accumulator settings
    hbm, tflops, src = peaks()
    ref = None if quick else load_reference()
    res = {"peaks": {"hbm_gbs": hbm, "bf16_tflops": tflops, "source": src}, "quantize": [], "gemm_exact": [], "gemm_tc": []}
    for n in (() if quick else (1 << 20, 1 << 24, 1 << 28)):
        x = torch.randn(n, device="cuda")
        out = torch.empty_like(x)
        ms = time_ms(lambda: native.quantize_(x, out, 4, 3))
        row = {"n": n, "ours_ms": round(ms, 4), "ours_GBps": round(8 * n / ms / 1e6, 1), "ours_frac_hbm_" + src: round(8 * n / ms / 1e6 / hbm, 3)}
        if ref is not None:
            y = x.clone()
            msr = time_ms(lambda: ref.float_quantize(y, 4, 3))
            row.update(ref_ms=round(msr, 4), ref_GBps=round(8 * n / msr / 1e6, 1), speedup=round(msr / ms, 2))
        res["quantize"].append(row)
        del x, out
    for n in (() if quick else (64, 1024, 4096)):
        a, b = torch.randn(n, n, device="cuda"), torch.randn(n, n, device="cuda")
        it = 5 if n <= 1024 else 2
        ms = time_ms(lambda: quant_gemm(a, b, man=3, exp=4, mode="exact"), iters=it, warm=1, flush=False)
        row = {"n": n, "ours_exact_ms": round(ms, 3), "ours_exact_GFLOPs": round(2 * n ** 3 / ms / 1e6, 1)}
        if ref is not None:
            msr = time_ms(lambda: ref.quant_gemm(a, b, man=3, exp=4), iters=it, warm=1, flush=False)
            row.update(ref_ms=round(msr, 3), ref_GFLOPs=round(2 * n ** 3 / msr / 1e6, 1), exact_speedup=round(msr / ms, 2))
        res["gemm_exact"].append(row)
    shapes = ((4096, 4096, 4096), (8192, 8192, 8192), (524288, 64, 576), (65536, 64, 576), (32768, 256, 2304)) if quick else \
        ((1024, 1024, 1024), (4096, 4096, 4096), (8192, 8192, 8192), (524288, 64, 576), (32768, 256, 2304))
    kinds = ((native.OPERAND_BF16, torch.bfloat16, "bf16"), (native.OPERAND_E4M3, torch.float8_e4m3fn, "e4m3"),
             (native.OPERAND_TF32, torch.float32, "tf32"))
    for (M, N, K) in shapes:
        for mode, dt, name in (kinds[:1] if quick else kinds):
            if M * K * (4 if name == "tf32" else 2) > (8 << 30):
                continue
            a = torch.randn(M, K, device="cuda").to(dt)
            bt = torch.randn(N, K, device="cuda").to(dt)
            out = torch.empty(M, N, device="cuda")
            accs = ((8, 23, False, 0), (5, 10, False, 0), (5, 10, False, 256)) if quick else \
                ((8, 23, False, 0), (5, 10, False, 0), (5, 10, False, 1024), (5, 10, False, 256), (4, 3, False, 128), (5, 2, True, 256))
            for (e, m, kahan, ks) in accs:
                if ks and ks > K:
                    continue
                for variant, vname in ((native.GEMM_1CTA, "1cta_128x128"), (native.GEMM_2CTA, "2cta_256x256")):
                    if variant == native.GEMM_2CTA and (e, m) == (8, 23):
                        continue
                    ms = time_ms(lambda: native.gemm_tc(a, bt, e, m, kahan=kahan, k_slice=ks, operand_mode=mode, out=out, variant=variant),
                                 iters=5, warm=2, flush=False)
                    tf = 2.0 * M * N * K / ms / 1e9
                    res["gemm_tc"].append({"M": M, "N": N, "K": K, "operands": name, "acc": [e, m], "kahan": kahan, "k_slice": ks or K,
                                           "variant": vname, "ms": round(ms, 4), "TFLOPs": round(tf, 1),
                                           "frac_bf16_peak_" + src: round(tf / tflops, 3)})
            if name == "bf16":
                ms = time_ms(lambda: torch.matmul(a, bt.t()), iters=5, warm=2, flush=False)
                res["gemm_tc"].append({"M": M, "N": N, "K": K, "operands": "bf16", "cublas_ms": round(ms, 4),
                                       "cublas_TFLOPs": round(2.0 * M * N * K / ms / 1e9, 1)})
            del a, bt, out
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
