#!/usr/bin/env python
"""Copies the measured numbers worth judging from gpurun_out/ (scratch) into profiles/ (tracked):
bench JSON lines, micro-benchmarks, all-reduce sweeps (with a roofline column), launch lists and
per-kernel ncu summaries (key raw metrics + the hottest source lines / stall reasons)."""
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")
NVLINK_MEASURED_GBS = 770.0      # per direction per GPU, B200_PROFILING.md
KEY_METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
               "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
               "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
               "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
               "smsp__inst_executed.sum", "smsp__cycles_active.avg", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
               "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
               "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
               "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
               "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
               "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def ncu_summary(rep, out_md):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return False
    H, U = rows[0], rows[1]
    lines = [f"# ncu summary: {os.path.basename(rep)}", "", "`ncu --set full --clock-control none --import-source on` on one B200; numbers under the profiler are for attribution only, never bench values.", ""]
    for r in rows[2:]:
        name = r[H.index("Kernel Name")]
        lines += [f"## {name[:160]}", "", "| metric | value | unit |", "|---|---|---|"]
        for m in KEY_METRICS:
            if m in H:
                lines.append(f"| {m} | {r[H.index(m)]} | {U[H.index(m)]} |")
        lines.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    hdrs = [i for i, r in enumerate(srows) if len(r) > 3 and r[0] == "Address"]
    if hdrs:
        hi = hdrs[0]
        SH = srows[hi]
        end = hdrs[1] - 1 if len(hdrs) > 1 else len(srows)
        try:
            cs, ci, cn = SH.index("Warp Stall Sampling (All Samples)"), SH.index("Source"), SH.index("Instructions Executed")
        except ValueError:
            cs = None
        if cs is not None:
            data = [r for r in srows[hi + 1:end] if len(r) == len(SH)]
            tot = sum(float(r[cs] or 0) for r in data) or 1.0
            top = sorted(((float(r[cs] or 0), r[ci].strip()[:90], r[cn]) for r in data), reverse=True)[:16]
            byop = {}
            for r in data:
                t = r[ci].strip().split()
                if not t:
                    continue
                op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
                byop[op] = byop.get(op, 0.0) + float(r[cs] or 0)
            lines += [f"## warp-stall samples by SASS instruction (first captured launch, {len(data)} instructions)", "",
                      "by opcode: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(byop.items(), key=lambda kv: -kv[1])[:12]), "",
                      "| % of samples | executed | instruction |", "|---|---|---|"]
            for sv, t, n in top:
                lines.append(f"| {100 * sv / tot:.1f} | {n} | `{t}` |")
            lines.append("")
    open(out_md, "w").write("\n".join(lines) + "\n")
    return True


def launch_summary(csv_path, out_md, title):
    rows = list(csv.reader(open(csv_path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
    if not hdr:
        return
    H = rows[hdr[0]]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = {}
    for r in rows[hdr[0] + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except (ValueError, IndexError):
            continue
        a = agg.setdefault(r[ki][:110], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values()) or 1
    lines = [f"# {title}", "", "`ncu --metrics gpu__time_duration.sum --clock-control none` over 3 eager (no CUDA graph) training steps; cold-cache, serialised: compare SHARES.", "",
             f"total {tot / 1e6:.2f} ms over {sum(v[0] for v in agg.values())} launches", "", "| share | launches | total us | kernel |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        lines.append(f"| {100 * v[1] / tot:.1f}% | {v[0]} | {v[1] / 1e3:.1f} | `{k}` |")
    open(out_md, "w").write("\n".join(lines) + "\n")


def sweep_md(js, out_md):
    d = json.load(open(js))
    W = d.get("world")
    rows = d.get("sweep", [])
    if not rows:
        return
    lines = [f"# low-bit all-reduce sweep, {W} x B200 ((4,3) + APS, 8-bit wire)", "",
             "Device-timed (CUDA events, max over ranks).  busbw = fp32 payload / t * 2(W-1)/W (nccl-tests convention).  Roofline for the fused kernel = slower of",
             "(wire bytes that must cross NVLink per GPU / 770 GB/s measured) and (local HBM traffic ~ 13 B/element / measured copy bandwidth).",
             ("Every timed call reduces a fresh copy of the gradients (finite values)." if any(k.endswith("finite_after") for r in rows for k in r) else
              "NOTE: taken with an in-place timing loop - the values overflow to Inf/NaN after a few dozen calls at W = 8, so the fused columns time the exceptional path (pessimistic); superseded by the later sweeps."), "",
             "| fp32 bytes | fused one-shot (best of packed / raw) us | fused two-shot us | fused NVLS us | fused Kahan us | NCCL fp32 all_reduce us | reference algorithm us | best fused busbw GB/s | NCCL busbw GB/s | reference busbw GB/s | fused roofline us | fused / roofline |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        hbm = 6650.0
    extra = []
    for r in rows:
        n = r["elems"]
        best = min(v for k, v in r.items() if k.startswith("fused_") and k.endswith("_us") and "kahan" not in k and "phases" not in k
                   and "fp32_auto" not in k and "nvls_fast" not in k)          # bit-exact low-bit variants only
        link_us = (2 * (W - 1) / W * n * 1) / (NVLINK_MEASURED_GBS * 1e3)          # two-shot wire bytes per GPU per direction
        hbm_us = (13.0 * n) / (hbm * 1e3)
        roof = max(link_us, hbm_us)
        bw = max(v for k, v in r.items() if k.startswith("fused_") and k.endswith("busbw_fp32_GBps"))
        extra.append((r["fp32_bytes"], r.get("fused_fp32_auto_us"), r.get("fused_nvls_fast_us"), r.get("nccl_fp32_allreduce_us")))
        lines.append(f"| {r['fp32_bytes']} | {min(r.get('fused_oneshot_us', 1e9), r.get('fused_raw_us', 1e9)) if ('fused_oneshot_us' in r or 'fused_raw_us' in r) else '-'} | {r.get('fused_twoshot_us', '-')} | {r.get('fused_nvls_us', '-')} | {r.get('fused_auto_kahan_us', '-')} | "
                     f"{r.get('nccl_fp32_allreduce_us', '-')} | {r.get('reference_algo_us', '-')} | {bw} | {r.get('nccl_fp32_busbw_GBps', '-')} | {r.get('reference_algo_busbw_fp32_GBps', '-')} | "
                     f"{roof:.2f} | {roof / best:.3f} |")
    if any(e[1] is not None for e in extra):
        lines += ["", "In-switch modes: plain fp32 sum of the (8,23) path (`multimem.ld_reduce.add.f32`, reference = NCCL all_reduce) and the opt-in,",
                  "NOT bit-exact fp8 in-switch sum (`algo=\"nvls_fast\"`, fp16 accumulator inside the NVSwitch).", "",
                  "| fp32 bytes | fused fp32 (8,23) us | NCCL fp32 all_reduce us | nvls_fast (4,3)+APS us |", "|---|---|---|---|"]
        lines += [f"| {b} | {f if f is not None else '-'} | {n if n is not None else '-'} | {q if q is not None else '-'} |" for (b, f, q, n) in extra]
    open(out_md, "w").write("\n".join(lines) + "\n")


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else "r2t*"        # which gpurun_out/ sub-directories (glob); round 1 used "trip*"
    os.makedirs(os.path.join(PR, "bench"), exist_ok=True)
    os.makedirs(os.path.join(PR, "ncu"), exist_ok=True)
    for p in glob.glob(os.path.join(GO, only, "bench*.json")):
        if os.path.getsize(p) > 10:
            shutil.copy(p, os.path.join(PR, "bench", os.path.basename(os.path.dirname(p)) + "_" + os.path.basename(p)))
    for p in glob.glob(os.path.join(GO, only, "microbench.json")) + glob.glob(os.path.join(GO, only, "bench_gemm.json")):
        if os.path.getsize(p) > 10:
            shutil.copy(p, os.path.join(PR, os.path.basename(os.path.dirname(p)) + "_" + os.path.basename(p)))
    for p in glob.glob(os.path.join(GO, only, "sweep.json")) + glob.glob(os.path.join(GO, only, "check.json")) + \
            glob.glob(os.path.join(GO, only, "check_sweep.json")) + glob.glob(os.path.join(GO, only, "fake8.json")) + \
            glob.glob(os.path.join(GO, only, "check_default.json")) + glob.glob(os.path.join(GO, only, "fake8_default.json")):
        tag = os.path.basename(os.path.dirname(p))
        shutil.copy(p, os.path.join(PR, f"allreduce_{tag}_{os.path.basename(p)}"))
        if p.endswith("sweep.json") or p.endswith("check_default.json"):
            sweep_md(p, os.path.join(PR, f"allreduce_{tag}_sweep.md"))
    for p in glob.glob(os.path.join(GO, only, "launches.csv")) + glob.glob(os.path.join(GO, only, "launches_smoke.csv")):
        launch_summary(p, os.path.join(PR, f"launches_{os.path.basename(os.path.dirname(p))}.md"), "kernel launch list of one ResNet-18 training step (ours)")
    for p in glob.glob(os.path.join(GO, only, "*.ncu-rep")):
        tag = os.path.basename(os.path.dirname(p)) + "_" + os.path.basename(p).replace(".ncu-rep", "")
        ok = ncu_summary(p, os.path.join(PR, "ncu", tag + ".md"))
        print("ncu", tag, ok)
    print("profiles updated")


if __name__ == "__main__":
    main()
