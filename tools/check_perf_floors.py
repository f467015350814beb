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
