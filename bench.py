#!/usr/bin/env python
"""Headline benchmark: training throughput (images/s) of the BASELINE.json configs.

    python bench.py --gpus N --steps K --warmup W [--model resnet18|resnet50|davidnet] [--impl ours|reference]

N > 1 is launched by the driver as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
Rank 0 prints ONE JSON line.  `value` is the whole-job images/s of the device-timed step
(CUDA events, barrier + synchronize on both sides, max over ranks); `e2e` is the same metric
through the public API (TrainEngine.train_step) with a pinned host->device copy of the inputs
and a device->host read of the loss every step.  Synthetic data of the named shape, random
init of the named architecture (no network on the box).
"""
from __future__ import annotations

import faulthandler

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs 2-4; batch sizes from the reference scripts (res18_cifar.yaml:4, main.py:32, dawn.py:15)
    "resnet18": dict(arch="resnet18_cifar", batch=512, shape=(3, 32, 32), classes=10, grad_exp=4, grad_man=3, use_aps=True,
                     use_kahan=False, use_lars=False, nesterov=False, lr=0.1, wd=1e-4, loss="mean", label="ResNet-18/CIFAR-shape"),
    "resnet50": dict(arch="resnet50", batch=32, shape=(3, 224, 224), classes=1000, grad_exp=5, grad_man=2, use_aps=True,
                     use_kahan=False, use_lars=False, nesterov=True, lr=0.1, wd=1e-4, loss="mean", label="ResNet-50/ImageNet-shape"),
    "davidnet": dict(arch="davidnet", batch=512, shape=(3, 32, 32), classes=10, grad_exp=5, grad_man=2, use_aps=True,
                     use_kahan=True, use_lars=True, nesterov=False, lr=0.1, wd=5e-4, loss="mean", label="DavidNet/CIFAR-shape"),
    # the custom-precision GEMM on a hot path: every convolution of ResNet-18 is a Quant_Conv with an fp16-shaped (5,10)
    # accumulator (reference: its own Quant_Conv = F.unfold + tvm_gemm; ours: tcgen05 tile mode, bf16 operands, fused
    # im2col/cast staging).  Per-GPU batch 64: the reference's SIMT GEMM needs ~1 s per step already.
    "resnet18_quant": dict(arch="resnet18_cifar_quant", batch=64, shape=(3, 32, 32), classes=10, grad_exp=4, grad_man=3, use_aps=True,
                           use_kahan=False, use_lars=False, nesterov=False, lr=0.1, wd=1e-4, loss="mean",
                           gemm_exp=5, gemm_man=10, label="ResNet-18/CIFAR-shape with Quant_Conv (5,10) GEMM accumulators"),
}
# Engine options of the "ours" arm and their defaults = what measured fastest on the B200 box (profiles/README.md);
# every one has a --no-... switch.  They change HOW the step is executed, never WHAT it computes (same architecture,
# fp32/TF32 compute, same batch, same gradient format).
# Measured on one B200, ResNet-18 config (profiles/bench/r2t3_*): stock BN + arena accumulation 69 230 img/s (e2e 67 884);
# fused BN 72 222; fused BN + stolen gradients + pipelined input copy 73 697 (e2e 73 287).  Overlapped bucketed reduction
# LOSES on this workload (1 GPU 7.57 vs 7.29 ms, 2 GPUs 7.53-7.84 vs 7.48 ms: the comm grid slows the convolutions it
# co-runs with by more than the ~0.1 ms of all-reduce it hides), so it stays off.
DEFAULT_FLAGS = dict(overlap=False, pipelined=True, fused_bn=True, steal_grads=True)

# published reference numbers: only ResNet-50 has one (<= ~1068 img/s on 8xV100, derived bound)
PUBLISHED = {"resnet50": 1068.0}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (rank 0 only)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.3] or [r for (_t, r) in self.rows[-3:]]
        sm = sorted(float(r[1]) for r in rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(rows)}


def shared_config(model_key: str, cfg: dict, world: int, emulate_node: int = 1) -> dict:
    """The part of `config` that names the benchmark: IDENTICAL keys and values in both arms (the driver compares them)."""
    out = {"model": cfg["arch"], "global_batch": cfg["batch"] * emulate_node * world, "per_gpu_batch": cfg["batch"],
           "image_shape": list(cfg["shape"]), "parallelism": f"dp{world}", "grad_format": f"exp{cfg['grad_exp']}man{cfg['grad_man']}",
           "use_APS": cfg["use_aps"], "use_kahan": cfg["use_kahan"], "use_lars": cfg["use_lars"], "emulate_node": emulate_node,
           "optimizer": "LARS" if cfg["use_lars"] else "SGD",
           "l2": "per-step working set (activations + gradients) exceeds the 126 MB L2; no explicit flush"}
    if "gemm_exp" in cfg:
        out["gemm_accumulator_format"] = f"exp{cfg['gemm_exp']}man{cfg['gemm_man']}"
    return out


METRIC = "{label} training images/sec (device-timed, max over ranks)"
DTYPE = "fp32 (cuDNN TF32 convolutions, torch defaults — the reference's stock fp32 path on both arms)"
DATA = "synthetic (random-init weights, random images/labels of the named shape)"


def timed(fn_step, steps: int, torch, dist, world: int):
    """barrier + sync | CUDA events around exactly `steps` calls | sync; max over ranks (ms)."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(steps):
        fn_step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, t0, t1


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from cpd_b200 import PrecisionConfig, dist_init
    from cpd_b200.data import PinnedBatchStream
    from cpd_b200.models import create_model
    from cpd_b200.optim import bn_bias_no_decay
    from cpd_b200.train import TrainEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = 0
    if world > 1:
        rank, world = dist_init()
    else:
        torch.cuda.set_device(0)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    pc = PrecisionConfig(grad_exp=cfg["grad_exp"], grad_man=cfg["grad_man"], use_aps=cfg["use_aps"], use_kahan=cfg["use_kahan"],
                         algo=args.algo, overlap=args.overlap, bucket_mb=args.bucket_mb, emulate_node=args.emulate_node,
                         comm_ctas=args.comm_ctas)
    model_kw = dict(exp=cfg["gemm_exp"], man=cfg["gemm_man"], mode="tile", operands="bf16") if "gemm_exp" in cfg else {}
    model_kw["fused_bn"] = bool(args.fused_bn)   # BatchNorm + shortcut add + ReLU in this repo's kernels (same arithmetic)
    model = create_model(cfg["arch"], **model_kw)
    wd_mult = bn_bias_no_decay(model.named_parameters()) if args.model == "resnet50" else None
    eng = TrainEngine(model, pc, batch_size=cfg["batch"], input_shape=cfg["shape"], lr=cfg["lr"], momentum=0.9,
                      weight_decay=cfg["wd"], use_lars=cfg["use_lars"], nesterov=cfg["nesterov"], wd_mult=wd_mult,
                      loss_reduction=cfg["loss"], channels_last=not args.nchw, cuda_graph=not args.no_graph,
                      fold_single=(args.model == "resnet50"), steal_grads=args.steal_grads)
    stream = PinnedBatchStream(cfg["batch"] * pc.emulate_node, cfg["shape"], cfg["classes"], pool=4, seed=1234 + rank)

    # warm-up through the public API (also captures the CUDA graph)
    losses = []
    for _ in range(max(args.warmup, 3) + 3):
        x, y = next(stream)
        losses.append(eng.train_step(x, y))
    eng.check()

    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    # (1) device-timed step: inputs already resident in the static buffers
    ms_dev, t0, _ = timed(eng.run_step, args.steps, torch, dist, world)

    # (2) end to end through the public API: pinned H2D of the inputs + D2H of the loss every step
    if args.pipelined:
        # optional: the next batch's pinned H2D copy runs on a copy stream behind the current step (still one H2D and
        # one loss D2H per step inside the timed region)
        eng.prefetch(*next(stream))

        def e2e_step():
            x, y = next(stream)
            losses.append(eng.train_step_pipelined(x, y))
    else:
        def e2e_step():
            x, y = next(stream)
            losses.append(eng.train_step(x, y))
    ms_e2e, _, t1 = timed(e2e_step, args.steps, torch, dist, world)
    clocks = sampler.stop(t0, t1) if sampler else None
    eng.check()

    imgs = cfg["batch"] * pc.emulate_node * world * args.steps
    value = imgs / (ms_dev / 1e3)
    e2e = imgs / (ms_e2e / 1e3)
    # counted by the native library while ONE step was issued (every launcher bumps an atomic); the captured CUDA graph
    # replays exactly these kernels every step, and the e2e loop runs the same graph
    launches_per_step = int(eng.kernel_launches_per_step or 0)
    if rank == 0:
        config = shared_config(args.model, cfg, world, pc.emulate_node)
        out = {
            "metric": METRIC.format(label=cfg["label"]), "value": round(value, 1),
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "warmup_untimed_steps_run": max(args.warmup, 3) + 3,      # W (>= 3) + 3 more so the CUDA-graph capture is never timed
            "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": (round(value / PUBLISHED[args.model], 3) if args.model in PUBLISHED else None),
            "dtype": DTYPE, "data": DATA, "impl": "ours", "config": config,
            "impl_detail": {"allreduce": eng.reducer.mode + (":" + (eng.reducer.symm.backend if eng.reducer.symm else "-")),
                            "nvls": bool(eng.reducer.symm.has_multicast) if eng.reducer.symm else False,
                            "wire_bits": eng.reducer.wire, "buckets": len(eng.reducer.buckets), "overlap": bool(pc.overlap), "comm_ctas": int(pc.comm_ctas), "bucket_mb": float(pc.bucket_mb),
                            "thresholds": eng.reducer.thresholds.to_dict(), "cuda_graph": not args.no_graph,
                            "channels_last": not args.nchw, "e2e_input_pipelining": bool(args.pipelined),
                            "steal_grads": bool(args.steal_grads), "fused_bn": bool(args.fused_bn)},
            "e2e": {"value": round(e2e, 1), "unit": "images/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                    "h2d_bytes_per_step": stream.bytes_per_batch, "d2h_bytes_per_step": 8,
                    "note": "TrainEngine.train_step" + ("_pipelined" if args.pipelined else "") +
                            ": pinned H2D of the batch + D2H of the loss and of the all-reduce status word, every step"},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_note": f"{launches_per_step} cpd_b200 kernel launches per step, counted by the native library "
                                 "(cpd_b200._C.launch_count) while the step was captured; the CUDA graph replays them every step",
            "clocks": clocks, "final_loss": round(losses[-1], 4),
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args, cfg):
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_harness
    ok, why = ref_harness.available()
    if not ok:
        print(json.dumps({"impl": "reference", "unavailable": why}))
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    ref_harness.setup_env(rank, world)
    try:
        import torch
        import torch.distributed as dist
        runner = ref_harness.RefRunner(args.model, cfg["batch"], cfg["grad_exp"], cfg["grad_man"], cfg["use_aps"],
                                       cfg["use_kahan"], cfg["use_lars"], gemm_exp=cfg.get("gemm_exp", 5), gemm_man=cfg.get("gemm_man", 10))
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {str(e)[:300]}"}))
        return
    losses = []
    for _ in range(max(args.warmup, 3)):
        losses.append(runner.step())
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ms, t0, t1 = timed(lambda: losses.append(runner.step()), args.steps, torch, dist, world)
    clocks = sampler.stop(t0, t1) if sampler else None
    imgs = cfg["batch"] * world * args.steps
    value = imgs / (ms / 1e3)
    if rank == 0:
        out = {
            "metric": METRIC.format(label=cfg["label"]), "value": round(value, 1),
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "warmup_untimed_steps_run": max(args.warmup, 3),
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": (round(value / PUBLISHED[args.model], 3) if args.model in PUBLISHED else None),
            "dtype": DTYPE, "data": DATA, "impl": "reference", "config": shared_config(args.model, cfg, world, 1),
            "impl_detail": {"path": "unmodified drcut/CPD from baseline/_ref: its models + CPDtorch.utils.dist_util.sum_gradients "
                                    "+ CPDtorch.quant.float_quantize (NCCL all_gather + per-parameter elementwise kernels)"
                                    + ("; every nn.Conv2d replaced by the reference's own Quant_Conv (F.unfold + tvm_gemm)"
                                       if "gemm_exp" in cfg else "")},
            "e2e": {"value": round(value, 1), "unit": "images/s", "h2d_bytes_per_step": cfg["batch"] * (4 * cfg["shape"][0] * cfg["shape"][1] * cfg["shape"][2] + 8),
                    "d2h_bytes_per_step": 4, "note": "the reference step already includes the pinned H2D copy and the loss .item()"},
            "gpu_launches": 0, "clocks": clocks, "final_loss": round(losses[-1], 4),
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_allreduce_sweep(args):
    """BASELINE.json config 5: low-bit all-reduce time / bus GB/s, 1 KB .. 1 GB of fp32 payload, fused variants vs plain
    NCCL fp32 all-reduce vs the reference algorithm (all_gather + W x {add, quantize}), device-timed, max over ranks."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import torch
    import torch.distributed as dist
    import multigpu_check as mg
    from cpd_b200 import dist_init
    rank, world = dist_init()
    res = {}
    mg.section_sweep(rank, world, res, args.sweep_max_bytes, quiet=True)
    if rank == 0:
        rows = res["sweep"]
        keep = ("fp32_bytes", "fused_auto_us", "fused_auto_busbw_fp32_GBps", "fused_auto_kahan_us", "fused_raw_us", "fused_oneshot_us",
                "fused_twoshot_us", "fused_nvls_us", "nccl_fp32_allreduce_us", "nccl_fp32_busbw_GBps", "reference_algo_us",
                "reference_algo_busbw_fp32_GBps", "fused_fp32_auto_us", "fused_fp32_auto_busbw_GBps", "fused_nvls_fast_us")
        big = max(rows, key=lambda r: r["fp32_bytes"])
        print(json.dumps({"metric": "low-bit all-reduce effective fp32 bus GB/s at the largest message ((4,3)+APS, fused, auto variant)",
                          "value": big.get("fused_auto_busbw_fp32_GBps"), "unit": "GB/s", "n_gpus": world, "higher_is_better": True,
                          "impl": "ours", "config": {"sweep": "1 KB .. %d B fp32 payload" % args.sweep_max_bytes, "format": "exp4man3+APS",
                                                     "timing": "CUDA events, max over ranks, fresh gradient buffer per call"},
                          "rows": [{k: r.get(k) for k in keep if k in r} for r in rows]}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


faulthandler.enable()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--model", choices=sorted(CONFIGS), default="resnet18")
    ap.add_argument("--algo", default="auto", choices=["auto", "oneshot", "twoshot", "nvls", "raw", "nvls_fast", "nccl"])
    ap.add_argument("--allreduce-sweep", dest="allreduce_sweep", action="store_true",
                    help="BASELINE config 5: all-reduce time / bus GB/s sweep instead of a training benchmark (needs --gpus >= 2)")
    ap.add_argument("--sweep-max-bytes", dest="sweep_max_bytes", type=int, default=1 << 30)
    ap.add_argument("--overlap", action=argparse.BooleanOptionalAction, default=DEFAULT_FLAGS["overlap"],
                    help="bucketed gradient reduction launched from autograd hooks on a comm stream, small comm grids")
    ap.add_argument("--bucket-mb", dest="bucket_mb", type=float, default=0.0)
    ap.add_argument("--comm-ctas", dest="comm_ctas", type=int, default=0, help="grid of an overlapped bucket's reduction (0: 32)")
    ap.add_argument("--emulate-node", dest="emulate_node", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--nchw", action="store_true")
    ap.add_argument("--fused-bn", dest="fused_bn", action=argparse.BooleanOptionalAction, default=DEFAULT_FLAGS["fused_bn"],
                    help="BatchNormAct2d: BatchNorm + shortcut add + ReLU fused (csrc/cuda/bn_act_sm100.cu) instead of cuDNN BN + ATen elementwise")
    ap.add_argument("--steal-grads", dest="steal_grads", action=argparse.BooleanOptionalAction, default=DEFAULT_FLAGS["steal_grads"],
                    help="backward hands over fresh gradient tensors; one multi-tensor copy per bucket moves them into the arena")
    ap.add_argument("--pipelined", action=argparse.BooleanOptionalAction, default=DEFAULT_FLAGS["pipelined"], help="e2e loop through prefetch / train_step_pipelined (H2D overlapped)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # convenience: self-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    cfg = CONFIGS[args.model]
    if args.allreduce_sweep:
        if int(os.environ.get("WORLD_SIZE", "1")) < 2:
            print(json.dumps({"metric": "allreduce sweep", "unavailable": "needs --gpus >= 2"}))
            return
        run_allreduce_sweep(args)
        return
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
