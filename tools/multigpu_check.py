#!/usr/bin/env python
"""Multi-GPU bring-up / correctness / bandwidth check of the fused all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/multigpu_check.py [--sections topo,symm,correct,stress,sweep] [--out gpurun_out/mg_N.json]

correct: fused kernel (one-shot / two-shot / NVLS, every wire width) == gather-mode reducer
         (NCCL all_gather + local fold kernel, itself verified against the C++/python oracles)
         bit for bit, on every rank, for several formats and layouts; all ranks bit-identical.
stress : thousands of back-to-back launches with alternating bucket sizes (flag/epoch/ABA races).
sweep  : time per call and bus bandwidth 1 KB .. 1 GB(fp32 payload) for fused variants, plain NCCL
         fp32 all_reduce and the reference algorithm (all_gather + W x {add, quantize}).
Device-timed with CUDA events, max over ranks.
"""
from __future__ import annotations

import faulthandler

import argparse
import json
import os
import sys
import time
import traceback

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpd_b200 import PrecisionConfig, dist_init  # noqa: E402
from cpd_b200.ops import native  # noqa: E402
from cpd_b200.parallel import ModelArena, GradReducer, discover_topology  # noqa: E402
from cpd_b200.parallel.arena import ArenaLayout  # noqa: E402


class FakeArena:
    """Just enough of ModelArena for GradReducer: a layout + a gradient arena."""

    def __init__(self, numels, device):
        self.layout = ArenaLayout(numels)
        self.device = device
        self.grad_flat = self.layout.new(device)
        self.model_flat = None
        self.params = []
        self._tables = {}

    def table(self, t0=0, t1=None):
        key = (t0, t1)
        if key not in self._tables:
            self._tables[key] = self.layout.table(self.device, t0, t1)
        return self._tables[key]

    bucket_ranges = ModelArena.bucket_ranges


def fill(arena, rank, seed, scale_lo=-20, scale_hi=2):
    g = torch.Generator(device="cpu").manual_seed(seed * 1000 + rank)
    flat = torch.zeros(arena.layout.total)
    for t, n in enumerate(arena.layout.numels):
        sc = 2.0 ** float(torch.randint(scale_lo, scale_hi, (1,), generator=g))
        flat[arena.layout.offsets[t]:arena.layout.offsets[t] + n] = torch.randn(n, generator=g) * sc
    arena.grad_flat.copy_(flat)


def max_over_ranks(ms):
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def time_calls(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return max_over_ranks(e0.elapsed_time(e1) / iters)


def time_calls_restored(fn, restore, iters, warm=2):
    """Per-call CUDA events around fn(); restore() runs between the calls, outside the timed region.  For large
    messages (launch latency negligible): the SAME buffer is reduced every time, like a training step's gradient
    arena (warm TLB), while the values stay those of the original gradients."""
    for _ in range(warm):
        restore()
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        restore()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return max_over_ranks(total / iters)


def same_bits(a, b):
    return bool(((a.view(torch.int32) == b.view(torch.int32)) | (a.isnan() & b.isnan())).all())


def all_ranks_equal(t, world):
    buf = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(buf, t)
    return all(same_bits(buf[0], b) for b in buf[1:])


SPECIALS = False      # --specials: add a layout whose tensors carry NaN / Inf / zeros / fp32 denormals on different ranks


def section_correct(rank, world, res):
    layouts = {"tiny": [1, 10, 64, 5], "mixed": [10, 64, 1, 4097, 300, 70000, 5, 128, 36864], "r18": None}
    from cpd_b200.models import create_model
    layouts["r18"] = [p.numel() for p in create_model("resnet18_cifar").parameters()]
    cases = [dict(grad_exp=4, grad_man=3, use_aps=True), dict(grad_exp=5, grad_man=2, use_aps=True, use_kahan=True),
             dict(grad_exp=5, grad_man=10, use_aps=True), dict(grad_exp=3, grad_man=0, use_aps=True),
             dict(grad_exp=5, grad_man=2, use_aps=False), dict(grad_exp=4, grad_man=3, use_aps=False, use_kahan=True),
             dict(grad_exp=4, grad_man=3, use_aps=False, prequantize=True), dict(grad_exp=8, grad_man=23, use_aps=False),
             dict(grad_exp=6, grad_man=9, use_aps=True, use_kahan=True)]
    out = []
    dev = torch.device("cuda", torch.cuda.current_device())
    if SPECIALS:
        layouts["specials"] = [4112, 1000, 64, 16, 7777, 300, 48]
    for lname, numels in layouts.items():
        arena = FakeArena(numels, dev)
        for ci, kw in enumerate(cases):
            oracle = GradReducer(arena, PrecisionConfig(**kw), mode="gather")
            fill(arena, rank, seed=ci * 7 + len(numels))
            if lname == "specials":
                # exceptional values, placed on different ranks: NaN, +Inf, an all-zero tensor, -Inf, fp32 denormals
                off, g = arena.layout.offsets, arena.grad_flat
                if rank == 0:
                    g[off[1] + 5] = float("nan")
                if rank == world - 1:
                    g[off[2] + 3] = float("inf")
                g[off[3]:off[3] + 16] = 0.0
                if rank == 1 % world:
                    g[off[4] + 11] = -float("inf")
                gen = torch.Generator().manual_seed(1000 * rank + ci)
                g[off[5]:off[5] + 300] = (torch.randn(300, generator=gen) * 1e-41).to(dev)
            mine = arena.grad_flat.clone()
            if kw.get("grad_exp") == 8 and kw.get("grad_man") == 23 and not kw.get("use_kahan") and not kw.get("use_aps"):
                # plain fp32: the gather-mode path is NCCL's all_reduce (its own summation order); the fused kernel sums in
                # rank order, so the oracle here is an explicit rank-ordered fp32 sum of the gathered gradients
                bufs = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(bufs, mine)
                want = torch.zeros_like(mine)
                for b_ in bufs:
                    want = want + b_
            else:
                oracle.reduce()
                want = arena.grad_flat.clone()
            for algo in ("oneshot", "twoshot", "nvls", "raw"):
                if algo == "raw" and (not kw.get("use_aps") or arena.layout.total > (1 << 22)):
                    continue
                fused = GradReducer(arena, PrecisionConfig(algo=algo, **kw), mode="fused")
                if algo == "nvls" and not fused.symm.has_multicast:
                    fused.close()
                    continue
                ok = True
                for rep in range(3):
                    arena.grad_flat.copy_(mine)
                    fused.reduce()
                    torch.cuda.synchronize()
                    fused.check_status()
                    ok = ok and same_bits(arena.grad_flat, want)
                ident = all_ranks_equal(arena.grad_flat, world)
                shifts_ok = bool(torch.equal(fused.shifts[0], oracle.shifts[0])) if kw.get("use_aps") else True
                out.append(dict(layout=lname, case=kw, algo=algo, wire=fused.wire, backend=fused.symm.backend, match=ok,
                                ranks_identical=ident, shifts=shifts_ok))
                fused.close()
                del fused
            del oracle
    res["correct"] = out
    res["correct_all_pass"] = all(o["match"] and o["ranks_identical"] and o["shifts"] for o in out)


def section_stress(rank, world, res, iters):
    dev = torch.device("cuda", torch.cuda.current_device())
    big = FakeArena([4097, 300, 70000, 36864, 128, 1 << 20], dev)
    cfg = PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, bucket_mb=0.2)      # several buckets of different sizes
    oracle = GradReducer(big, cfg, mode="gather")
    out = {}
    for algo in ("oneshot", "twoshot", "nvls"):
        fused = GradReducer(big, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, bucket_mb=0.2, algo=algo), mode="fused")
        if algo == "nvls" and not fused.symm.has_multicast:
            fused.close()
            continue
        fill(big, rank, seed=99)
        mine = big.grad_flat.clone()
        oracle.reduce()
        want = big.grad_flat.clone()
        bad = 0
        t0 = time.time()
        for i in range(iters):
            big.grad_flat.copy_(mine)
            fused.reduce()
            if i % 97 == 0:
                torch.cuda.synchronize()
                bad += 0 if same_bits(big.grad_flat, want) else 1
        torch.cuda.synchronize()
        fused.check_status()
        bad += 0 if same_bits(big.grad_flat, want) else 1
        out[algo] = dict(iters=iters, buckets=len(fused.buckets), mismatching_checks=bad, seconds=round(time.time() - t0, 2))
        fused.close()
        del fused
    res["stress"] = out
    res["stress_all_pass"] = all(v["mismatching_checks"] == 0 for v in out.values())


def section_fault(rank, world, res):
    """rank `world - 1` reaches the collective late (torch.cuda._sleep, the same hook as
    CPD_B200_FAULT_DELAY_RANK); with a short timeout every OTHER rank must return (no hang), poison its gradients and
    raise a RuntimeError that names the late rank."""
    from cpd_b200.parallel import reducer as R
    dev = torch.device("cuda", torch.cuda.current_device())
    late = world - 1
    out = {}
    for algo in ("oneshot", "twoshot"):
        arena = FakeArena([4096, 300, 70000], dev)
        fill(arena, rank, seed=11)
        red = GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, algo=algo), mode="fused",
                          timeout_cycles=20_000_000)                               # ~10 ms
        red.reduce()                                                                # a healthy launch first
        torch.cuda.synchronize()
        red.check_status()
        dist.barrier()
        old = (R._FAULT_DELAY_RANK, R._FAULT_DELAY_CYCLES)
        R._FAULT_DELAY_RANK, R._FAULT_DELAY_CYCLES = late, 4e8                      # ~200 ms >> timeout
        fill(arena, rank, seed=12)
        red.reduce()
        torch.cuda.synchronize()
        R._FAULT_DELAY_RANK, R._FAULT_DELAY_CYCLES = old
        msg = ""
        try:
            red.check_status()
        except RuntimeError as e:
            msg = str(e)
        ok = (rank == late) or (f"waiting for rank {late}" in msg and bool(torch.isnan(arena.grad_flat[:4096]).all()))
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        out[algo] = dict(all_on_time_ranks_raised_and_poisoned=bool(t.item()), rank0_error=msg if rank == 0 else None)
        red.close()
        del red
        dist.barrier()
    res["fault"] = out
    res["fault_all_pass"] = all(v["all_on_time_ranks_raised_and_poisoned"] for v in out.values())


def reference_algorithm(g, world, e, m, gather_buf):
    """dist_util.py:54-69 with our quantize kernel: all_gather fp32 + W x {add, quantize}."""
    bufs = [gather_buf[r] for r in range(world)]
    dist.all_gather(bufs, g)
    res = torch.zeros_like(g)
    for b in bufs:
        res = res + b
        native.quantize_(res, res, e, m)
    g.copy_(res)


def section_sweep(rank, world, res, max_bytes, quiet=False):
    dev = torch.device("cuda", torch.cuda.current_device())
    rows = []
    sizes = [1 << k for k in range(10, 31, 2) if (1 << k) <= max_bytes]       # fp32 payload bytes
    for nbytes in sizes:
        n = nbytes // 4
        arena = FakeArena([n], dev)
        fill(arena, rank, seed=3, scale_lo=-8, scale_hi=-2)
        iters = 200 if nbytes <= (1 << 20) else (40 if nbytes <= (1 << 26) else 8)
        row = dict(fp32_bytes=nbytes, elems=n)
        # Every timed call reduces a FRESH copy of the gradients (a pool of iters + warm-up buffers, refilled before
        # each variant): an in-place loop feeds each result back as the next input, the values grow ~W x per call,
        # overflow to Inf/NaN after a few dozen calls at W = 8, and the kernels then time their exceptional paths.
        # The pool is also larger than L2 for all but the smallest sizes, so the inputs are cold like a real step's.
        K = iters + 3
        orig = arena.grad_flat.clone()
        pool = orig.unsqueeze(0).repeat(K, 1)
        cursor = [0]

        def refill():
            pool.copy_(orig.unsqueeze(0).expand(K, -1))
            cursor[0] = 0

        def next_buf():
            b = pool[cursor[0] % K]
            cursor[0] += 1
            return b

        for algo in ("raw", "oneshot", "twoshot", "nvls"):
            if algo == "oneshot" and nbytes > (1 << 26):
                continue
            if algo == "raw" and nbytes > (1 << 22):
                continue
            red = GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, algo=algo), mode="fused")
            if algo == "nvls" and not red.symm.has_multicast:
                red.close()
                continue
            refill()
            ms = time_calls(lambda: red.reduce(next_buf()), iters)
            row[f"fused_{algo}_us"] = round(ms * 1e3, 2)
            row[f"fused_{algo}_phases_us"] = red.phase_times_us()
            row[f"fused_{algo}_busbw_fp32_GBps"] = round(nbytes / (ms * 1e-3) * 2 * (world - 1) / world / 1e9, 2)
            row[f"fused_{algo}_wire_GBps"] = round((n * red.wire / 8) / (ms * 1e-3) * 2 * (world - 1) / world / 1e9, 2)
            row[f"fused_{algo}_finite_after"] = bool(torch.isfinite(pool).all())
            if nbytes >= (1 << 26) and algo in ("twoshot", "nvls"):
                # training-like access pattern: one arena, reduced in place every step, values restored off the clock
                g = arena.grad_flat
                ms = time_calls_restored(lambda: red.reduce(), lambda: g.copy_(orig), max(iters // 2, 4))
                row[f"fused_{algo}_same_buffer_us"] = round(ms * 1e3, 2)
            red.close()
            del red
        for kname, kalgo in (("auto", "auto"),):
            redk = GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, use_kahan=True, algo=kalgo), mode="fused")
            refill()
            ms = time_calls(lambda: redk.reduce(next_buf()), iters)
            row[f"fused_{kname}_kahan_us"] = round(ms * 1e3, 2)
            redp = GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, algo=kalgo), mode="fused")
            refill()
            ms = time_calls(lambda: redp.reduce(next_buf()), iters)
            row[f"fused_{kname}_us"] = round(ms * 1e3, 2)
            row[f"fused_{kname}_busbw_fp32_GBps"] = round(nbytes / (ms * 1e-3) * 2 * (world - 1) / world / 1e9, 2)
            redk.close(); redp.close()
            del redk, redp
        # plain fp32 sum: the (8,23) path.  Reference = NCCL all_reduce (dist_util.py:55-59); ours = in-switch multimem.ld_reduce
        redf = GradReducer(arena, PrecisionConfig(grad_exp=8, grad_man=23), mode="fused")
        refill()
        ms = time_calls(lambda: redf.reduce(next_buf()), iters)
        row["fused_fp32_auto_us"] = round(ms * 1e3, 2)
        row["fused_fp32_auto_algo"] = int(redf.algo_for(redf.buckets[0]))
        row["fused_fp32_auto_busbw_GBps"] = round(nbytes / (ms * 1e-3) * 2 * (world - 1) / world / 1e9, 2)
        redf.close()
        del redf
        # opt-in, NOT bit-exact: fp8 codes summed inside the switch with an fp16 accumulator
        redq = GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, algo="nvls_fast"), mode="fused")
        if redq.symm.has_multicast and nbytes >= (1 << 16):
            refill()
            ms = time_calls(lambda: redq.reduce(next_buf()), iters)
            row["fused_nvls_fast_us"] = round(ms * 1e3, 2)
        redq.close()
        del redq
        refill()
        ms = time_calls(lambda: dist.all_reduce(next_buf()), iters)
        row["nccl_fp32_allreduce_us"] = round(ms * 1e3, 2)
        if nbytes >= (1 << 26):
            g = arena.grad_flat
            ms2 = time_calls_restored(lambda: dist.all_reduce(g), lambda: g.copy_(orig), max(iters // 2, 4))
            row["nccl_fp32_same_buffer_us"] = round(ms2 * 1e3, 2)
        row["nccl_fp32_busbw_GBps"] = round(nbytes / (ms * 1e-3) * 2 * (world - 1) / world / 1e9, 2)
        if nbytes <= (1 << 28):
            gb = torch.empty(world, orig.numel(), device=dev)
            refill()
            ms = time_calls(lambda: reference_algorithm(next_buf(), world, 4, 3, gb), max(iters // 4, 3))
            row["reference_algo_us"] = round(ms * 1e3, 2)
            row["reference_algo_busbw_fp32_GBps"] = round(nbytes / (ms * 1e-3) * 2 * (world - 1) / world / 1e9, 2)
            del gb
        del pool, orig
        rows.append(row)
        if rank == 0 and not quiet:
            print(json.dumps(row), flush=True)
        del arena
        torch.cuda.empty_cache()
    res["sweep"] = rows


faulthandler.enable()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sections", default="topo,symm,correct,stress,sweep")
    ap.add_argument("--out", default="")
    ap.add_argument("--stress-iters", type=int, default=2000)
    ap.add_argument("--max-bytes", type=int, default=1 << 30)
    ap.add_argument("--specials", action="store_true", help="correct section: also a layout with NaN / Inf / zero / denormal tensors")
    args = ap.parse_args()
    global SPECIALS
    SPECIALS = bool(args.specials)
    rank, world = dist_init()
    res = {"world": world, "torch": torch.__version__}
    secs = args.sections.split(",")
    try:
        if "topo" in secs:
            topo = discover_topology()
            res["topology"] = topo.summary()
            res["topology_raw"] = {k: topo.raw.get(k) for k in ("device_count", "nvlink", "multicast", "cuda_driver", "cuda_runtime")}
            if rank == 0:
                print("[topo]", topo.summary(), flush=True)
        if "symm" in secs:
            from cpd_b200.parallel.symm import SymmetricWorkspace
            st = {}
            for b in ("vmm", "torch", "ipc"):
                try:
                    ws = SymmetricWorkspace(1 << 20, 64, backend=b)
                    st[b] = dict(ok=True, multicast=ws.has_multicast, log=ws.log)
                    del ws
                except Exception as e:  # noqa: BLE001
                    st[b] = dict(ok=False, error=f"{type(e).__name__}: {str(e)[:300]}")
            res["symm_backends"] = st
            if rank == 0:
                print("[symm]", json.dumps(st), flush=True)
        if "correct" in secs:
            section_correct(rank, world, res)
            if rank == 0:
                print("[correct] all_pass =", res["correct_all_pass"], "cases =", len(res["correct"]), flush=True)
                for o in res["correct"]:
                    if not (o["match"] and o["ranks_identical"] and o["shifts"]):
                        print("   FAIL", o, flush=True)
        if "stress" in secs:
            section_stress(rank, world, res, args.stress_iters)
            if rank == 0:
                print("[stress]", json.dumps(res["stress"]), flush=True)
        if "sweep" in secs:
            section_sweep(rank, world, res, args.max_bytes)
        if "fault" in secs:                       # last: the timed-out reducers leave their status words set
            section_fault(rank, world, res)
            if rank == 0:
                print("[fault]", json.dumps(res["fault"]), flush=True)
    except Exception:  # noqa: BLE001
        res["error"] = traceback.format_exc()
        print(f"[rank {rank}] ERROR\n{res['error']}", flush=True)
    if rank == 0 and args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)
    ok = res.get("correct_all_pass", True) and res.get("stress_all_pass", True) and res.get("fault_all_pass", True) and "error" not in res
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
