"""Trainer-side library features: schedules, samplers, meters, checkpoint/resume, config, façade."""
import argparse
import math
import os

import numpy as np
import pytest
import torch

from cpd_b200 import PrecisionConfig
from cpd_b200.config import add_precision_args
from cpd_b200.data import (DistributedGivenIterationSampler, GivenIterationSampler, DistributedSampler, Crop, FlipLR, Cutout, Transform,
                           cifar_shape, PinnedBatchStream, normalise, pad, transpose)
from cpd_b200.optim import WarmupStepLR, PiecewiseLinear, IterLRScheduler, linear_warmup
from cpd_b200.utils import AverageMeter, accuracy, save_checkpoint, load_state, make_state, find_latest, JsonlLogger, TableLogger


def test_resnet18_schedule_matches_recipe():
    it_per_epoch = 12
    s = WarmupStepLR(0.1, 1.6, 5 * it_per_epoch, [40 * it_per_epoch, 80 * it_per_epoch])
    assert s(0) == pytest.approx(0.1) and s(5 * it_per_epoch) == pytest.approx(1.6)
    assert s(30) == pytest.approx(0.1 + 1.5 * 30 / 60)
    assert s(40 * it_per_epoch) == pytest.approx(0.16) and s(80 * it_per_epoch + 1) == pytest.approx(0.016)


def test_piecewise_linear_and_iter_scheduler():
    p = PiecewiseLinear([0, 5, 24], [0, 0.4, 0])
    assert p(0) == 0 and p(5) == pytest.approx(0.4) and p(24) == 0 and p(2.5) == pytest.approx(0.2)
    it = IterLRScheduler(0.1, [10, 20], [0.1, 0.5])
    assert it(5) == pytest.approx(0.1) and it(10) == pytest.approx(0.01) and it(25) == pytest.approx(0.005)
    assert [it.step() for _ in range(3)] == [pytest.approx(0.1)] * 3
    assert linear_warmup(1.0, 5, 10) == 0.5 and linear_warmup(1.0, 11, 10) == 1.0


def test_given_iteration_samplers_partition_and_resume():
    ds = list(range(103))
    full = [list(DistributedGivenIterationSampler(ds, 7, 4, world_size=3, rank=r)) for r in range(3)]
    assert all(len(x) == 28 for x in full)
    a = DistributedGivenIterationSampler(ds, 7, 4, world_size=3, rank=1, last_iter=2)
    assert list(a) == full[1][12:]
    with pytest.raises(RuntimeError):
        list(a)
    g = list(GivenIterationSampler(ds, 5, 8))
    assert len(g) == 40 and set(g) <= set(range(103))
    s0, s1 = DistributedSampler(ds, world_size=2, rank=0), DistributedSampler(ds, world_size=2, rank=1)
    assert len(list(s0)) == len(list(s1)) == 52 and set(list(s0)) | set(list(s1)) == set(range(103))
    s0.set_epoch(1)
    assert list(s0) != list(DistributedSampler(ds, world_size=2, rank=0))


def test_augmentations_and_preprocessing():
    raw = np.random.randint(0, 255, (10, 32, 32, 3)).astype(np.uint8)
    x = transpose(normalise(pad(raw, 4)))
    assert x.shape == (10, 3, 40, 40) and x.dtype == np.float32
    t = Transform(list(zip(x, range(10))), [Crop(32, 32), FlipLR(), Cutout(8, 8)])
    t.set_random_choices()
    img, label = t[3]
    assert img.shape == (3, 32, 32) and label == 3 and (img == 0).sum() >= 3 * 64
    ds = cifar_shape(64)
    assert ds[0][0].shape == (3, 32, 32) and len(ds) == 64
    st = PinnedBatchStream(4, pin=False)
    assert next(st)[0].shape == (4, 3, 32, 32) and st.bytes_per_batch == 4 * 3 * 32 * 32 * 4 + 4 * 8


def test_meters_and_accuracy():
    m = AverageMeter(3)
    for v in (1, 2, 3, 4):
        m.update(v)
    assert m.avg == 3 and m.val == 4
    g = AverageMeter()
    g.update(2, 2); g.update(4, 2)
    assert g.avg == 3
    out = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.1], [0.2, 0.3, 0.5]])
    top1, top2 = accuracy(out, torch.tensor([1, 1, 2]), topk=(1, 2))
    assert top1.item() == pytest.approx(200 / 3) and top2.item() == pytest.approx(100.0)


def test_checkpoint_save_resume_reproduces_next_step(tmp_path):
    from cpd_b200.train import TrainEngine
    torch.manual_seed(0)
    def make():
        torch.manual_seed(0)
        m = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(48, 16), torch.nn.ReLU(), torch.nn.Linear(16, 10))
        return TrainEngine(m, PrecisionConfig(grad_exp=5, grad_man=2, use_aps=True), batch_size=4, input_shape=(3, 4, 4), lr=0.05,
                           use_lars=False, cuda_graph=False)
    x, y = torch.randn(4, 3, 4, 4), torch.randint(0, 10, (4,))
    a = make()
    for _ in range(3):
        a.train_step(x, y)
    path = str(tmp_path / "ckpt_3.pth")
    save_checkpoint(make_state(a.model, a.optimizer, step=3, precision=a.cfg), path, is_best=True)
    nxt = a.train_step(x, y)
    b = make()
    best, last = load_state(path, b.model, b.optimizer)
    assert last == 3 and os.path.exists(str(tmp_path / "ckpt_best.pth"))
    assert b.train_step(x, y) == pytest.approx(nxt, rel=1e-6)
    assert torch.equal(a.arena.model_flat, b.arena.model_flat)
    assert find_latest(str(tmp_path / "ckpt_*.pth")) == path
    # module. prefix fix-up
    sd = {"state_dict": {"module." + k: v for k, v in a.model.state_dict().items()}, "step": 9}
    torch.save(sd, str(tmp_path / "m.pth"))
    assert load_state(str(tmp_path / "m.pth"), b.model)[1] == 9


def test_model_only_and_foreign_optimizer_checkpoints_survive_a_step(tmp_path):
    """a checkpoint without a cpd optimizer entry must not be reverted by the next step (the step copies
    the fp32 masters into the model), and a reference-format torch.optim.SGD state_dict must load, not KeyError."""
    from cpd_b200.train import TrainEngine
    def make():
        torch.manual_seed(0)
        m = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(48, 16), torch.nn.ReLU(), torch.nn.Linear(16, 10))
        return TrainEngine(m, PrecisionConfig(grad_exp=8, grad_man=23), batch_size=4, input_shape=(3, 4, 4), lr=0.0,
                           weight_decay=0.0, use_lars=False, cuda_graph=False)
    x, y = torch.randn(4, 3, 4, 4), torch.randint(0, 10, (4,))
    eng = make()
    sd = {"state_dict": {k: torch.full_like(v, 0.5) for k, v in eng.model.state_dict().items()}, "step": 1}
    torch.save(sd, str(tmp_path / "model_only.pth"))
    load_state(str(tmp_path / "model_only.pth"), eng.model, eng.optimizer)
    eng.train_step(x, y)                                   # lr = 0: the weights must stay what was loaded
    assert all(bool((p == 0.5).all()) for p in eng.model.parameters())
    # reference-style checkpoint: torch SGD state_dict under 'optimizer' (mix.py:345-356)
    ref_model = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(48, 16), torch.nn.ReLU(), torch.nn.Linear(16, 10))
    opt = torch.optim.SGD(ref_model.parameters(), lr=0.3, momentum=0.9)
    ref_model(x).sum().backward()
    opt.step()
    torch.save({"state_dict": ref_model.state_dict(), "optimizer": opt.state_dict(), "step": 2}, str(tmp_path / "ref.pth"))
    eng2 = make()
    _, last = load_state(str(tmp_path / "ref.pth"), eng2.model, eng2.optimizer)
    assert last == 2 and eng2.optimizer.lr == pytest.approx(0.3)
    for p, q in zip(eng2.model.parameters(), ref_model.parameters()):
        assert torch.equal(p.detach(), q.detach())
    moms = [eng2.arena.layout.view(eng2.optimizer.mom, t, like=p.data) for t, p in enumerate(eng2.arena.params)]
    for mview, q in zip(moms, ref_model.parameters()):
        assert torch.equal(mview, opt.state[q]["momentum_buffer"])
    eng2.set_lr(0.0)
    eng2.optimizer.momentum = 0.0
    eng2.train_step(x, y)
    for p, q in zip(eng2.model.parameters(), ref_model.parameters()):
        assert torch.equal(p.detach(), q.detach())


def test_config_flags_accept_both_spellings():
    p = argparse.ArgumentParser()
    add_precision_args(p)
    a = p.parse_args(["--use-APS", "--grad_exp", "4", "--grad-man", "3", "--emulate-node", "8", "--use_kahan"])
    c = PrecisionConfig.from_args(a)
    assert (c.grad_exp, c.grad_man, c.use_aps, c.use_kahan, c.emulate_node) == (4, 3, True, True, 8)
    b = p.parse_args(["--use_APS", "--emulate_node", "2"])
    assert PrecisionConfig.from_args(b).emulate_node == 2
    with pytest.raises(ValueError):
        PrecisionConfig(grad_exp=9)


def test_cpdtorch_facade_imports():
    from CPDtorch.quant import float_quantize, quantizer, Quantizer, quant_gemm, Quant_Linear, Quant_Conv  # noqa: F401
    from CPDtorch.utils.dist_util import dist_init, DistModule, sum_gradients, normal_sum_gradients, kahan_sum_gradients, broadcast_params  # noqa: F401
    assert float_quantize(torch.tensor([1.06]), 4, 3).item() == 1.0
    # the reference's module paths work too (quant_function.py / quant_module.py)
    from CPDtorch.quant.quant_function import float_quantize as fq2, quant_gemm as g2, quantizer as q2  # noqa: F401
    from CPDtorch.quant.quant_module import Quantizer as Q2, Quant_Linear as L2, Quant_Conv as C2  # noqa: F401
    assert fq2 is float_quantize and L2 is Quant_Linear


def test_loggers(tmp_path, capsys):
    j = JsonlLogger(str(tmp_path / "log.jsonl"))
    j.log(step=1, loss=0.5)
    j.close()
    assert '"loss": 0.5' in open(tmp_path / "log.jsonl").read()
    t = TableLogger()
    t.append({"epoch": 1, "loss": 0.25})
    assert "epoch" in capsys.readouterr().out


@pytest.mark.parametrize("E", [4, 32])          # 32: the reference README's "8 GPUs emulate 256 nodes" setting
def test_emulate_node_equals_multi_rank_fold(E):
    """emulate_node=E on one process == the W=E fold of the same micro-gradients (bitwise)."""
    from cpd_b200.parallel import ModelArena, EmulateNode
    from cpd_b200.ops import native
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(12, 9), torch.nn.Tanh(), torch.nn.Linear(9, 3))
    arena = ModelArena(list(net.parameters()))
    cfg = PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, emulate_node=E)
    em = EmulateNode(arena, cfg)
    xs = [torch.randn(5, 12) for _ in range(E)]
    micro = []
    for k in range(E):
        em.begin_micro(k)
        (net(xs[k]).square().mean() / E).backward()
        micro.append(em.slots[k].clone())
    em.fold()
    want = torch.empty(arena.layout.total)
    native.aps_fold(micro, want, arena.layout.table("cpu").seg_off, 4, 3, True, False, float(E))
    assert torch.equal(arena.grad_flat, want)
    assert net[0].weight.grad.data_ptr() == arena.grad_views[0].data_ptr()


def test_aps_optimizer_hook_runner_interface():
    from types import SimpleNamespace
    from cpd_b200.parallel import APSOptimizerHook
    torch.manual_seed(0)
    model = torch.nn.Linear(6, 3)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    hook = APSOptimizerHook(grad_exp=5, grad_man=2, use_APS=True, grad_clip={"max_norm": 1.0})
    x, y = torch.randn(8, 6), torch.randn(8, 3)
    w0 = model.weight.detach().clone()
    runner = SimpleNamespace(model=model, optimizer=opt, outputs={"loss": torch.nn.functional.mse_loss(model(x), y)})
    hook.after_train_iter(runner)
    assert not torch.equal(model.weight.detach(), w0)
    from cpd_b200 import float_quantize
    g = model.weight.grad
    assert torch.isfinite(g).all()


def test_graphnet_from_nested_dict_matches_builtin_davidnet():
    import torch.nn as nn
    from cpd_b200.models import davidnet, graph_from_nested, rel, Add, Scale, Identity, Correct

    def conv_bn(cin, cout):
        return {"conv": nn.Conv2d(cin, cout, 3, padding=1, bias=False), "bn": nn.BatchNorm2d(cout), "relu": nn.ReLU(True)}

    def residual(c):
        return {"in": Identity(), "res1": conv_bn(c, c), "res2": conv_bn(c, c), "add": (Add(), [rel("in"), rel("res2", "relu")])}

    spec = {"prep": conv_bn(3, 64),
            "layer1": dict(conv_bn(64, 128), pool=nn.MaxPool2d(2), residual=residual(128)),
            "layer2": dict(conv_bn(128, 256), pool=nn.MaxPool2d(2)),
            "layer3": dict(conv_bn(256, 512), pool=nn.MaxPool2d(2), residual=residual(512)),
            "classifier": {"pool": nn.MaxPool2d(4), "flatten": nn.Flatten(1), "linear": nn.Linear(512, 10, bias=False), "logits": Scale(0.125)},
            "correct": (Correct(), [("classifier", "logits"), "target"])}
    g = graph_from_nested({k: v for k, v in spec.items() if k != "correct"}, ["classifier_logits"])
    ref = davidnet()
    assert sum(p.numel() for p in g.parameters()) == sum(p.numel() for p in ref.parameters()) == 6573120
    assert len(list(g.parameters())) == 25
    g.load_state_dict({k2: v for (k2, _), v in zip(g.state_dict().items(), ref.state_dict().values())})
    x = torch.randn(2, 3, 32, 32)
    g.eval(); ref.eval()
    assert torch.allclose(g(x), ref(x), atol=1e-5)


def test_comm_stats_and_profiler_trace(tmp_path):
    """Observability: APS shift histogram + guard counters for the JSONL log, chrome trace export."""
    import json
    from cpd_b200.train import TrainEngine
    from cpd_b200.train.loop import LoopConfig, run
    from cpd_b200.utils import trace, nvtx_range
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(48, 16), torch.nn.ReLU(), torch.nn.Linear(16, 10))
    eng = TrainEngine(m, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True), batch_size=4, input_shape=(3, 4, 4), lr=0.05,
                      use_lars=False, cuda_graph=False)
    x, y = torch.randn(4, 3, 4, 4), torch.randint(0, 10, (4,))
    eng.train_step(x, y)
    st = eng.comm_stats()
    assert sum(st["aps_shift_hist"].values()) == 4 and st["aps_shift_min"] <= st["aps_shift_max"]      # 4 parameter tensors
    log = str(tmp_path / "log.jsonl")
    tr = str(tmp_path / "trace.json")
    run(eng, [(x, y)] * 6, LoopConfig(max_iter=6, lr_fn=lambda s: 0.05, print_freq=2, log_jsonl=log, profile_trace=tr, profile_start=2, profile_steps=2))
    rows = [json.loads(line) for line in open(log)]
    assert rows and "aps_shift_hist" in rows[-1] and "images_per_sec" in rows[-1]
    assert os.path.getsize(tr) > 100 and "traceEvents" in open(tr).read(4096)
    with trace(None) as p, nvtx_range("noop"):
        assert p is None


def test_committed_profiles_hold_the_perf_floors():
    """The evidence under profiles/ must keep satisfying tools/check_perf_floors.py (SURVEY §4 'Perf regression')."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "check_perf_floors.py")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-3000:]


def test_engine_autocast_compute_keeps_fp32_gradient_path():
    """dawn.py --half / --bf16: forward/backward under autocast, parameters + gradients + all-reduce path stay fp32."""
    from cpd_b200.train import TrainEngine
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(48, 32), torch.nn.ReLU(), torch.nn.Linear(32, 10))
    eng = TrainEngine(m, PrecisionConfig(grad_exp=5, grad_man=2, use_aps=True), batch_size=8, input_shape=(3, 4, 4), lr=0.05,
                      use_lars=False, cuda_graph=False, amp_dtype=torch.bfloat16)
    x, y = torch.randn(8, 3, 4, 4), torch.randint(0, 10, (8,))
    first = eng.train_step(x, y)
    for _ in range(20):
        last = eng.train_step(x, y)
    assert last < first and eng.arena.grad_flat.dtype == torch.float32 and all(p.dtype == torch.float32 for p in m.parameters())
    out, loss = eng.evaluate(x, y)
    assert out.dtype == torch.float32 and torch.isfinite(loss)


def test_pipelined_input_prefetch_equals_plain_steps():
    """prefetch / train_step_pipelined (H2D of step i+1 behind the compute of step i) == the same batches through train_step."""
    from cpd_b200.train import TrainEngine

    def make():
        torch.manual_seed(0)
        m = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(48, 16), torch.nn.ReLU(), torch.nn.Linear(16, 10))
        return TrainEngine(m, PrecisionConfig(grad_exp=5, grad_man=2, use_aps=True), batch_size=4, input_shape=(3, 4, 4), lr=0.05,
                           use_lars=False, cuda_graph=False)
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(4, 3, 4, 4, generator=g), torch.randint(0, 10, (4,), generator=g)) for _ in range(5)]
    a, b = make(), make()
    plain = [a.train_step(x, y) for x, y in batches]
    b.prefetch(*batches[0])
    piped = [b.train_step_pipelined(*(batches[i + 1] if i + 1 < len(batches) else (None, None))) for i in range(len(batches))]
    assert piped == plain and torch.equal(a.arena.model_flat, b.arena.model_flat)
    with pytest.raises(RuntimeError):
        b.begin_step()
