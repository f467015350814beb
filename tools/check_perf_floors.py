#!/usr/bin/env python
"""Performance-regression floors (SURVEY §4 'Perf regression').  Reads the benchmark JSON that the GPU trips leave
under gpurun_out/ (or, by default, the copies committed under profiles/) and asserts fraction-of-roofline /
versus-baseline floors, so a change that silently loses a kernel's speed fails a check instead of a review.

    python tools/check_perf_floors.py            # committed evidence in profiles/
    python tools/check_perf_floors.py --dir gpurun_out/trip18_n8 ...   # fresh results
Exit code 1 and one line per violated floor."""
import argparse
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest(pattern):
    def trip_no(p):
        m = re.search(r"trip(\d+)", p)
        return int(m.group(1)) if m else -1
    files = sorted(glob.glob(pattern), key=trip_no)
    return files[-1] if files else None


def _bench_line(path):
    for line in reversed(open(path).read().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    return None


def checks(prof):
    out = []          # (name, value, floor, ok)

    def floor(name, value, lo):
        out.append((name, value, lo, value is not None and value >= lo))

    g = _latest(os.path.join(prof, "trip*_bench_gemm.json"))
    if g:
        d = json.load(open(g))
        hbm = d["peaks"]["hbm_gbs"]
        q = [r for r in d.get("quantize", []) if r["n"] >= (1 << 28)]
        if q:
            floor("quantize 256Mi: fraction of measured HBM copy bandwidth", q[-1]["ours_GBps"] / hbm, 0.80)
            floor("quantize 256Mi: speed-up over the reference kernel", q[-1]["speedup"], 3.5)
        ge = [r for r in d.get("gemm_exact", []) if r["n"] == 4096]
        if ge:
            floor("gemm_exact 4096^3: speed-up over the reference tvm_gemm", ge[-1]["exact_speedup"], 3.0)
        tc = [r for r in d.get("gemm_tc", []) if r.get("M") == 8192]
        cublas = [r["cublas_TFLOPs"] for r in tc if "cublas_TFLOPs" in r]
        if cublas:
            def best(operands, k_slice, variant):
                v = [r["TFLOPs"] for r in tc if r.get("operands") == operands and r.get("k_slice") == k_slice and r.get("variant") == variant
                     and r.get("acc") == [5, 10] and not r.get("kahan")]
                return max(v) if v else None
            floor("tcgen05 2-CTA GEMM 8192^3 bf16, (5,10) requant every 1024: fraction of cuBLAS bf16", (best("bf16", 1024, "2cta_256x256") or 0) / cublas[0], 0.70)
            floor("tcgen05 2-CTA GEMM 8192^3 bf16, (5,10) requant every 256: fraction of cuBLAS bf16", (best("bf16", 256, "2cta_256x256") or 0) / cublas[0], 0.60)
            floor("tcgen05 2-CTA GEMM 8192^3 e4m3, (5,10) requant every 1024: multiple of cuBLAS bf16", (best("e4m3", 1024, "2cta_256x256") or 0) / cublas[0], 1.15)
    mx = _latest(os.path.join(prof, "bench", "trip*_bench_mx.json"))
    if mx:
        rows = [r for r in json.load(open(mx)) if r["n"] == 8192 and r["acc"] == [8, 23]]
        if rows:
            floor("block-scaled (MX) fp8 GEMM 8192^3: TFLOP/s", rows[-1]["TFLOPs"], 1000.0)
    sw = _latest(os.path.join(prof, "allreduce_trip*_n8_check_sweep.json"))
    if sw:
        d = json.load(open(sw))
        floor("8-GPU all-reduce: every correctness case bit-identical", 1.0 if d.get("correct_all_pass") else 0.0, 1.0)
        for r in d["sweep"]:
            fused = min(v for k, v in r.items() if re.fullmatch(r"fused_(raw|oneshot|twoshot|nvls)_us", k))
            nccl, nb = r["nccl_fp32_allreduce_us"], r["fp32_bytes"]
            if nb <= 4096 or nb >= (4 << 20):
                floor(f"8-GPU all-reduce {nb} B: NCCL fp32 time / fused time", nccl / fused, 1.10)
            else:
                floor(f"8-GPU all-reduce {nb} B (mid band, open item): NCCL fp32 time / fused time", nccl / fused, 0.80)
            if "reference_algo_us" in r:
                floor(f"8-GPU all-reduce {nb} B: reference algorithm time / fused time", r["reference_algo_us"] / fused, 4.0)
            if nb >= (1 << 30):
                floor("8-GPU all-reduce 1 GB: fp32-equivalent bus bandwidth GB/s", max(v for k, v in r.items() if k.endswith("busbw_fp32_GBps") and k.startswith("fused")), 850.0)
    b1, b8 = _latest(os.path.join(prof, "bench", "trip*_bench_ours.json")), _latest(os.path.join(prof, "bench", "trip*_n8_bench_ours.json"))
    one = [p for p in glob.glob(os.path.join(prof, "bench", "trip*_bench_ours.json")) if "_n" not in os.path.basename(p)]
    ref8 = _latest(os.path.join(prof, "bench", "trip*_n8_bench_ref.json"))
    if one and b8:
        j1, j8 = _bench_line(sorted(one)[-1]), _bench_line(b8)
        if j1 and j8:
            floor("ResNet-18 weak scaling 1 -> 8 GPUs (ms/step ratio)", j1["ms_per_step"] / j8["ms_per_step"], 0.95)
            floor("ResNet-18 8 GPUs: end-to-end / device-timed throughput", j8["e2e"]["value"] / j8["value"], 0.95)
            if ref8 and _bench_line(ref8):
                floor("ResNet-18 8 GPUs: ours / unmodified reference (images/s)", j8["value"] / _bench_line(ref8)["value"], 5.0)
    del b1
    # ---- round 2 evidence -------------------------------------------------------------------------------------------
    sw2 = [p for p in sorted(glob.glob(os.path.join(prof, "allreduce_r2t*_n8_check_default.json")))
           if "stress_all_pass" in json.load(open(p))]                  # (the final sweep-only report is checked further down)
    if sw2:
        d = json.load(open(sw2[-1]))
        floor("r2 8-GPU all-reduce: correctness + stress + fault injection all pass",
              1.0 if (d.get("correct_all_pass") and d.get("stress_all_pass") and d.get("fault_all_pass")) else 0.0, 1.0)
        for r in d["sweep"]:
            nb, auto, nccl = r["fp32_bytes"], r["fused_auto_us"], r["nccl_fp32_allreduce_us"]
            # at or ahead of NCCL's fp32 all-reduce at EVERY size (NCCL's own small-message numbers scatter ~10 %)
            floor(f"r2 8-GPU all-reduce {nb} B: NCCL fp32 time / fused AUTO time", nccl / auto, 1.0 if nb <= (1 << 20) else 1.4)
            if "reference_algo_us" in r and r["reference_algo_us"]:
                floor(f"r2 8-GPU all-reduce {nb} B: reference algorithm time / fused AUTO time", r["reference_algo_us"] / auto, 4.0)
    mb = sorted(glob.glob(os.path.join(prof, "r2t*_microbench.json")))
    if mb:
        d = json.load(open(mb[-1]))
        e2e = [r for r in d.get("gemm_tile_e2e", []) if r["n"] == 8192]
        if e2e:
            floor("r2 quant_gemm(fp32, fp32, tile) 8192^3: end-to-end / kernel-only throughput", e2e[-1]["e2e_over_kernel_only"], 0.70)
        f1 = [r for r in d.get("aps_fold", []) if r["W"] == 1 and not r["kahan"]]
        if f1:
            floor("r2 aps_fold_local W=1 on ResNet-18's gradients: ms (ceiling, as 1/ms)", 1.0 / f1[-1]["ms"], 1.0 / 0.050)
    def r2bench(name):
        c = sorted(glob.glob(os.path.join(prof, "bench", name)))
        return _bench_line(c[-1]) if c else None
    o1, r1 = r2bench("r2t4_bench_ours_x.json"), r2bench("r2t4_bench_ref_resnet18.json")
    if o1 and r1:
        floor("r2 ResNet-18 1 GPU: ours e2e / unmodified reference (images/s)", o1["e2e"]["value"] / r1["value"], 4.7)
        floor("r2 ResNet-18 1 GPU: end-to-end / device-timed throughput", o1["e2e"]["value"] / o1["value"], 0.98)
    q1, qr = r2bench("r2t4_bench_ours_xmodelresnet18_quantsteps15.json"), r2bench("r2t1_bench_quant_ref.json")
    if q1 and qr:
        floor("r2 ResNet-18 with Quant_Conv (5,10) accumulators, 1 GPU: ours / reference's own Quant_Conv", q1["value"] / qr["value"], 300.0)
    o8 = r2bench("r2t5_n8_bench_ours_resnet18.json")
    if o8 and o1:
        floor("r2 ResNet-18 weak scaling 1 -> 8 GPUs (ms/step ratio; both with the shipped defaults)", 6.768 / o8["ms_per_step"], 0.96)
    # ---- final tree of round 2 (trips 9-14) ---------------------------------------------------------------------------
    def r2file(rel):
        p = os.path.join(prof, rel)
        return _bench_line(p) if os.path.isfile(p) else None
    refs = {"resnet18": r2bench("r2t4_bench_ref_resnet18.json"), "resnet50": r2bench("r2t4_bench_ref_resnet50.json"),
            "davidnet": r2bench("r2t4_bench_ref_davidnet.json"), "resnet18_quant": qr}
    lo_ratio = {"resnet18": 5.0, "resnet50": 7.5, "davidnet": 2.8, "resnet18_quant": 1200.0}
    for cfg in ("resnet18", "resnet50", "davidnet", "resnet18_quant"):
        f1, f8 = r2file(f"r2/t14_bench_{cfg}_n1.json"), r2bench(f"r2t13_n8_bench_ours_{cfg}.json")
        if f1 and refs.get(cfg):
            floor(f"r2 final {cfg} 1 GPU: ours e2e / unmodified reference", f1["e2e"]["value"] / refs[cfg]["value"], lo_ratio[cfg])
        if f1 and f8:
            floor(f"r2 final {cfg}: weak-scaling efficiency 1 -> 8 GPUs (device-timed)", f8["value"] / (8 * f1["value"]), 0.94)
            floor(f"r2 final {cfg} 8 GPUs: end-to-end / device-timed throughput", f8["e2e"]["value"] / f8["value"], 0.97)
    g9 = os.path.join(prof, "r2", "t9_bench_gemm_persistent_raster.json")
    if os.path.isfile(g9):
        tc = [r for r in json.load(open(g9))["gemm_tc"] if r.get("M") == 8192]
        cub = [r["cublas_TFLOPs"] for r in tc if "cublas_TFLOPs" in r]
        pair = [r["TFLOPs"] for r in tc if r.get("variant") == "2cta_256x256" and r.get("k_slice") == 8192]
        one = [r["TFLOPs"] for r in tc if r.get("variant") == "1cta_128x128" and r.get("k_slice") == 8192 and r.get("acc") == [5, 10]]
        if cub and pair and one:
            floor("r2 final tcgen05 pair GEMM 8192^3 bf16, (5,10) cast at the end: fraction of cuBLAS bf16", pair[0] / cub[0], 0.85)
            floor("r2 final tcgen05 128x128 GEMM 8192^3 bf16: fraction of cuBLAS bf16", one[0] / cub[0], 0.55)
    qs, qi = r2file("r2/t9_bench_quant_staged.json"), r2file("r2/t14_bench_resnet18_quant_n1.json")
    if qs and qi:
        floor("r2 final Quant_Conv ResNet-18: implicit-GEMM step / staged-im2col step (throughput ratio)", qi["value"] / qs["value"], 1.3)
    for n, name in ((8, "allreduce_r2t13_n8_check_default.json"), (4, "allreduce_r2t5_n4_check_default.json")):
        pth = os.path.join(prof, name)
        if not os.path.isfile(pth):
            continue
        d = json.load(open(pth))
        floor(f"r2 final {n}-GPU all-reduce: every correctness case passes", 1.0 if d.get("correct_all_pass") else 0.0, 1.0)
        for r in d["sweep"]:
            nb, auto, nccl = r["fp32_bytes"], r["fused_auto_us"], r["nccl_fp32_allreduce_us"]
            lo = (1.0 if nb <= (1 << 20) else 1.4) if n == 8 else (0.9 if nb <= (1 << 20) else 1.35)
            floor(f"r2 final {n}-GPU all-reduce {nb} B: NCCL fp32 time / fused AUTO time", nccl / auto, lo)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default=os.path.join(ROOT, "profiles"))
    args = ap.parse_args()
    res = checks(args.dir)
    bad = [r for r in res if not r[3]]
    for name, value, lo, ok in res:
        print(f"{'ok  ' if ok else 'FAIL'} {name}: {value:.3f} (floor {lo})" if value is not None else f"FAIL {name}: no data")
    print(f"{len(res) - len(bad)}/{len(res)} floors hold")
    return 1 if bad or not res else 0


if __name__ == "__main__":
    sys.exit(main())
