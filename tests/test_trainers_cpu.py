"""The three example trainers (the reference's mix.py / dawn.py / main.py workloads) run end to end for a handful of
synthetic steps on the CPU backend: argument parsing incl. the legacy flag spellings, YAML config, samplers, LR
schedules, the TrainEngine step with APS / Kahan / LARS / emulate_node, validation, checkpoint + auto-resume, logs."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, cwd, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="4")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "SLURM_PROCID"):
        env.pop(k, None)
    r = subprocess.run([sys.executable] + args, capture_output=True, text=True, timeout=timeout, cwd=cwd, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.timeout(900)
def test_resnet18_mix_trainer_runs_validates_checkpoints_and_resumes(tmp_path):
    save = str(tmp_path / "r18")
    base = [os.path.join(ROOT, "examples/ResNet18/mix.py"), "--cpu", "--grad_exp", "4", "--grad_man", "3", "--use_APS",
            "--emulate_node", "2", "--batch_size", "4", "--dataset_len", "64", "--val-len", "16", "--workers", "0",
            "--print_freq", "2", "--val_freq", "5", "--save_path", save]
    out = _run(base + ["--max-iter", "10"], str(tmp_path))
    assert "Iter: [10/10]" in out and out.count(" * All Loss") == 2
    assert os.path.isfile(os.path.join(save, "ckpt_10.pth")) and os.path.isfile(os.path.join(save, "ckpt_5.pth"))
    rows = [json.loads(ln) for ln in open(os.path.join(save, "train.jsonl"))]
    train_rows = [r for r in rows if "loss" in r]
    assert all(r["loss"] == r["loss"] for r in train_rows) and any("aps_shift_hist" in r for r in train_rows)
    # resume from the step-10 checkpoint and continue with LARS
    out2 = _run(base + ["--max-iter", "12", "--load-path", os.path.join(save, "ckpt_10.pth"), "--use_lars"], str(tmp_path))
    assert "=> resumed from" in out2 and "at iteration 10" in out2 and "Iter: [12/12]" in out2
    # evaluation-only entry point
    out3 = _run(base + ["-e", "--load-path", os.path.join(save, "ckpt_10.pth")], str(tmp_path))
    assert " * All Loss" in out3


@pytest.mark.timeout(900)
def test_davidnet_dawn_trainer_runs_with_aps_kahan_lars(tmp_path):
    out = _run([os.path.join(ROOT, "examples/DavidNet/dawn.py"), "--cpu", "--dist", "0", "--epoch", "2", "-b", "8", "--dataset_len",
                "40", "--grad_exp", "5", "--grad_man", "2", "--use_APS", "--use_kahan", "--use_lars", "--warm_up_epoch", "1"],
               str(tmp_path))
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert any("train loss" in ln for ln in lines) and "nan" not in out.lower()
    assert any(ln.startswith("epoch\t") or "epoch" in ln for ln in lines)


@pytest.mark.timeout(900)
def test_resnet50_main_trainer_runs_and_auto_resumes(tmp_path):
    fmt = str(tmp_path / "ck" / "checkpoint-{epoch}.pth.tar")
    base = [os.path.join(ROOT, "examples/ResNet50/main.py"), "--no-cuda", "--arch", "resnet18", "--batch-size", "2", "--dataset-len", "8",
            "--val-len", "4", "--val-batch-size", "2", "--workers", "0", "--epochs", "2", "--use-APS", "--grad_exp", "5", "--grad_man", "2",
            "--emulate-node", "2", "--log-dir", str(tmp_path / "logs"), "--checkpoint-format", fmt]
    out = _run(base + ["--max-iter", "4"], str(tmp_path))
    assert "Iter: [1/4]" in out and out.count(" * All Loss") == 2       # 2 iterations per epoch: validation at 2 and 4
    assert os.path.isfile(str(tmp_path / "ck" / "ckpt_4.pth"))
    out2 = _run(base + ["--max-iter", "6"], str(tmp_path))            # resume="auto": picks up the newest checkpoint
    assert "=> resumed from" in out2 and "at iteration 4" in out2 and " * All Loss" in out2
    assert os.path.isfile(str(tmp_path / "ck" / "ckpt_6.pth"))


@pytest.mark.timeout(900)
def test_fcn_segmentation_trainer_runs_through_the_optimizer_hook(tmp_path):
    """The reference's FCN/Cityscapes workload (mmcv-fork optimizer hook, README.md:132-151): runner + APSOptimizerHook."""
    out = _run([os.path.join(ROOT, "examples/FCN/train_fcn.py"), "--cpu", "--arch", "fcn_r18_d8", "--width", "8", "--crop", "65",
                "--batch-size", "2", "--max-iters", "6", "--dataset-len", "8", "--val-len", "4", "--val-freq", "3", "--print-freq", "2",
                "--workers", "0", "--use_APS", "--use_kahan", "--grad-clip", "5"], str(tmp_path))
    assert "Iter [6/6]" in out and out.count(" * Iter") == 2 and "nan" not in out.lower()
    assert "lr 0.00207" in out                                   # poly schedule: (0.01 - 1e-4) * (1 - 5/6)^0.9 + 1e-4


def test_fcn_d8_topology():
    import torch
    from cpd_b200.models import create_model
    m = create_model("fcn_r50_d8")
    assert sum(p.numel() for p in m.parameters()) == 35_320_678       # ResNet-50 backbone + FCN head (512) + aux head (256), 19 classes
    small = create_model("fcn_r18_d8", width=8).eval()
    feats = []
    hooks = [st.register_forward_hook(lambda _m, _i, o: feats.append(o.shape[-1])) for st in small.stages]
    with torch.no_grad():
        y = small(torch.randn(1, 3, 129, 129))
    for h in hooks:
        h.remove()
    assert y.shape == (1, 19, 129, 129)
    assert feats == [33, 17, 17, 17]                                   # output stride 8: stages 3 and 4 dilate instead of striding


@pytest.mark.timeout(900)
def test_fcn_trainer_two_gloo_ranks(tmp_path):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = _run(["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
                os.path.join(ROOT, "examples/FCN/train_fcn.py"), "--cpu", "--arch", "fcn_r18_d8", "--width", "8", "--crop", "65",
                "--batch-size", "2", "--max-iters", "4", "--dataset-len", "8", "--val-len", "4", "--val-freq", "4", "--print-freq", "2",
                "--workers", "0", "--use_APS"], str(tmp_path))
    assert "batch 2 x 2" in out and "Iter [4/4]" in out and " * Iter 4: mIoU" in out and "nan" not in out.lower()
