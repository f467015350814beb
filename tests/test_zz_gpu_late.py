"""GPU tests written after round 1's GPU budget was spent: they have passed their CPU dry-runs but have not executed on
a GPU yet.  The file name sorts last on purpose, so that under ``pytest -x`` a problem in one of them cannot hide the
results of the suites that have a GPU track record."""
import os

import numpy as np
import pytest
import torch

from cpd_b200 import float_quantize, PrecisionConfig
from cpd_b200.ops import native
from test_gpu_kernels import _arenas, _same

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [8, 4])
def test_fused_allreduce_fake_world_matches_cpu_fold(world):
    """The cross-GPU kernel on ONE GPU: it is told world_size = W while every peer pointer aliases local memory
    (tools/fake_world.py), so the raw and packed one-shot variants run their full W-rank instruction stream —
    amax, exponent mailboxes, scale + cast + pack, W-rank decode + fold (plain and Kahan), unscale — and the result
    must equal the pure-software CPU fold of W copies bit for bit, incl. tensors with NaN / Inf / zeros / fp32 denormals."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29700 + world + os.getpid() % 200))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fake_world.py"), "--check-only", "--world", str(world)],
                       capture_output=True, text=True, timeout=280, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "mismatching cases: none" in r.stdout, r.stdout[-3000:]


@pytest.mark.timeout(200)
def test_fused_allreduce_timeout_poisons_outputs_and_names_the_missing_rank():
    """SURVEY §5.3 fault injection: a peer that never signals must not hang the kernel NOR leak stale data into the
    gradients — bounded spin, NaN-poisoned outputs, status word -> RuntimeError naming rank and phase (tools/fault_timeout.py)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29650 + os.getpid() % 200))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fault_timeout.py")], capture_output=True, text=True,
                       timeout=180, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("FAULT_TIMEOUT ")][-1]
    res = json.loads(line[len("FAULT_TIMEOUT "):])
    assert [x["algo"] for x in res] == ["raw", "oneshot", "twoshot"]
    for x in res:
        assert x["returned_ms"] < 1000, x                      # bounded: a few timeouts of ~2 ms, not a hang
        assert x["all_nan"], x
        assert "timed out" in x["error"] and "waiting for rank 1" in x["error"], x


@pytest.mark.parametrize("e,m", [(4, 3), (5, 2), (5, 10), (6, 9)])
@pytest.mark.parametrize("aps,kahan", [(True, False), (True, True), (False, False), (False, True)])
def test_aps_fold_32_sources_matches_cpu(e, m, aps, kahan):
    """emulate_node 32 (the reference README's 8-GPUs-emulate-256-nodes setting): the fold kernel takes up to 64 arenas."""
    W = 32
    numels = [10, 64, 1, 4097, 300, 70000, 5]
    layout, cpu = _arenas(W, numels, W * 7 + e, "cpu")
    tab_c, tab_g = layout.table("cpu"), layout.table("cuda")
    want = torch.empty(layout.total)
    sh_c = torch.zeros(len(numels), dtype=torch.int32)
    native.aps_fold(cpu, want, tab_c.seg_off, e, m, aps, kahan, float(W), shifts_out=sh_c)
    gpu = [a.cuda() for a in cpu]
    got = torch.empty(layout.total, device="cuda")
    ws = torch.zeros(4 + 2 * 64, dtype=torch.int32, device="cuda")
    sh_g = torch.zeros(len(numels), dtype=torch.int32, device="cuda")
    for rep in range(2):
        native.aps_fold(gpu, got, tab_g.seg_off, e, m, aps, kahan, float(W), chunks=tab_g.chunks, workspace=ws, shifts_out=sh_g)
        assert _same(got, want), (rep, e, m, aps, kahan)
        if aps:
            assert torch.equal(sh_g.cpu(), sh_c)


def test_quantize_more_than_2_31_elements():
    """64-bit indexing (the reference truncates the size to int, quant.cu:17): 2^31 + 4099 elements in place."""
    n = (1 << 31) + 4099
    free, _ = torch.cuda.mem_get_info()
    if free < n * 4 + (1 << 30):
        pytest.skip("not enough free device memory")
    x = torch.empty(n, device="cuda")
    x.fill_(1.2345)
    x[-4099:] = torch.linspace(-3.0, 3.0, 4099, device="cuda")
    tail = x[-4099:].clone().cpu()
    float_quantize(x, 4, 3, inplace=True)
    assert _same(x[-4099:], float_quantize(tail, 4, 3))
    assert _same(x[:1024], float_quantize(torch.full((1024,), 1.2345), 4, 3))
    # the 1024 elements that end exactly at the 2^31 boundary (the linspace tail starts there)
    mid = (1 << 31) - 1024
    assert _same(x[mid:mid + 1024], float_quantize(torch.full((1024,), 1.2345), 4, 3))
    # ... and a window that straddles 2^31: constant part | first 512 elements of the tail
    assert _same(x[(1 << 31) - 512:(1 << 31) + 512],
                 float_quantize(torch.cat([torch.full((512,), 1.2345), tail[:512]]), 4, 3))
    del x
    torch.cuda.empty_cache()



def test_overlapped_bucketed_reduce_equals_sequential():
    """buckets reduced on the comm stream as soon as their gradients are ready (hooks, eager and captured
    in a CUDA graph) give the same weights as reducing the same buckets after backward."""
    from cpd_b200.train import TrainEngine

    def run(overlap, graph, steal=False):
        torch.manual_seed(0)
        layers = []
        for _ in range(6):
            layers += [torch.nn.Linear(256, 256), torch.nn.ReLU()]
        m = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 8 * 8, 256), torch.nn.ReLU(), *layers, torch.nn.Linear(256, 10))
        eng = TrainEngine(m, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, bucket_mb=0.3, overlap=overlap), batch_size=64,
                          input_shape=(3, 8, 8), lr=0.05, use_lars=False, cuda_graph=graph, warmup_steps=1, channels_last=False,
                          steal_grads=steal)
        assert len(eng.reducer.buckets) >= 4
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(64, 3, 8, 8, generator=g).pin_memory(), torch.randint(0, 10, (64,), generator=g).pin_memory()
        losses = [eng.train_step(x, y) for _ in range(5)]
        torch.cuda.synchronize()
        # direct evidence that EVERY bucket was reduced: each gradient, times its tensor's APS scale, is a (4,3) member
        k = 0
        for b, sh in zip(eng.reducer.buckets, eng.reducer.shifts):
            for t in range(b.T):
                g = eng.arena.grad_views[b.t0 + t].detach().flatten().float() * (2.0 ** int(sh[t]))
                assert torch.equal(float_quantize(g, 4, 3), g), (overlap, graph, b.t0 + t)
                k += 1
        assert k == len(eng.arena.params)
        return eng.arena.model_flat.clone(), losses

    for graph in (False, True):
        w_seq, l_seq = run(False, graph)
        w_ovl, l_ovl = run(True, graph)
        # same kernels, same order per bucket -> equal; the criterion tolerates a stray rounding flip but not a bucket that
        # missed its reduce (at world size 1 that is the (4,3) quantisation of every gradient element: percent-level)
        bad = ((w_ovl - w_seq).abs() > 3e-8 + 1e-6 * w_seq.abs()).float().mean().item()      # an unreduced run: ~0.3
        assert bad < 1e-3, (graph, bad, float((w_ovl - w_seq).abs().max()))
        assert np.allclose(l_ovl, l_seq, rtol=1e-3), (graph, l_ovl, l_seq)
        assert l_seq[-1] < l_seq[0]
        # steal_grads: fresh gradient tensors + one multi-tensor copy per bucket, with and without overlap
        for ovl in (False, True):
            w_st, l_st = run(ovl, graph, steal=True)
            bad = ((w_st - w_seq).abs() > 3e-8 + 1e-6 * w_seq.abs()).float().mean().item()
            assert bad < 1e-3, (graph, ovl, bad)
            assert np.allclose(l_st, l_seq, rtol=1e-3), (graph, ovl, l_st, l_seq)


def test_pipelined_input_prefetch_equals_plain_steps_on_gpu():
    """prefetch / train_step_pipelined (pinned H2D of step i+1 on a copy stream behind the compute of step i, graph-captured
    step) == the same batches through train_step: same losses, same weights."""
    from cpd_b200.train import TrainEngine

    def make():
        torch.manual_seed(0)
        m = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 8 * 8, 128), torch.nn.ReLU(), torch.nn.Linear(128, 10))
        return TrainEngine(m, PrecisionConfig(grad_exp=5, grad_man=2, use_aps=True), batch_size=32, input_shape=(3, 8, 8), lr=0.05,
                           cuda_graph=True, warmup_steps=1, channels_last=False)
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(32, 3, 8, 8, generator=g).pin_memory(), torch.randint(0, 10, (32,), generator=g).pin_memory()) for _ in range(7)]
    a, b = make(), make()
    plain = [a.train_step(x, y) for (x, y) in batches]
    b.prefetch(*batches[0])
    piped = [b.train_step_pipelined(*(batches[i + 1] if i + 1 < len(batches) else (None, None))) for i in range(len(batches))]
    torch.cuda.synchronize()
    assert plain == piped, (plain, piped)
    assert torch.equal(a.arena.model_flat, b.arena.model_flat)


def test_implicit_gemm_tf32_operands():
    from test_gpu_implicit_conv import _operands
    g = (2, 32, 10, 10, 48, 3, 1, 1, 1)                     # 32 fp32 channels = one 128-byte block per tap
    B, C, H, W, O, k, stride, pad, dil = g
    x, w, _, _, (OH, OW) = _operands(g, torch.float32)
    # operands through the staging kernel itself (it rounds fp32 to tf32 the way it does for the explicit matrix)
    x_nhwc = native.stage_matrix(x.permute(0, 2, 3, 1).reshape(-1, C), native.OPERAND_TF32).view(B, H, W, C)
    wt = native.stage_matrix(w.permute(0, 2, 3, 1).reshape(O, -1), native.OPERAND_TF32)
    cols = native.stage_im2col(x, k, k, (OH, OW), stride, pad, dil, native.OPERAND_TF32)
    ref = native.gemm_tc(cols, wt, 8, 23, False, 0, native.OPERAND_TF32, None, variant=native.GEMM_1CTA, k=k * k * C)
    got = native.gemm_conv(x_nhwc, wt, k, k, (OH, OW), stride, pad, dil, 8, 23, False, 0, native.OPERAND_TF32, variant=native.GEMM_1CTA)
    assert torch.equal(got, ref)


def test_fused_bn_survives_a_change_of_batch_size():
    """The last batch of an epoch is smaller: another slab geometry over the SAME workspace.  The arrival counters live at
    a fixed offset, so stale partial sums of the bigger launch can never be read as counters (they could, before)."""
    import torch.nn.functional as F
    from cpd_b200.nn import BatchNormAct2d
    torch.manual_seed(0)
    C = 64
    fused = BatchNormAct2d(C, relu=True).cuda()
    stock = torch.nn.BatchNorm2d(C).cuda()
    stock.load_state_dict(fused.state_dict())
    for B in (64, 7, 64, 3, 33):
        x = torch.randn(B, C, 16, 16, device="cuda").contiguous(memory_format=torch.channels_last)
        xf, xs = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        y, ys = fused(xf), F.relu(stock(xs))
        assert torch.allclose(y, ys, rtol=1e-4, atol=1e-4), (B, float((y - ys).abs().max()))
        dy = torch.randn_like(ys)
        y.backward(dy); ys.backward(dy)
        assert torch.allclose(xf.grad, xs.grad, rtol=1e-3, atol=2e-4 * float(xs.grad.abs().max())), B
        assert torch.allclose(fused.running_mean, stock.running_mean, rtol=1e-5, atol=1e-5)
        assert torch.allclose(fused.running_var, stock.running_var, rtol=1e-4, atol=1e-5)


def test_fused_bn_statistics_are_bitwise_reproducible():
    """Two-level reduction of the slab partials: whichever CTAs arrive last, every sum runs in a fixed order."""
    from cpd_b200.nn import BatchNormAct2d
    torch.manual_seed(1)
    x = torch.randn(128, 64, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last)
    dy = torch.randn_like(x)
    outs = []
    for _ in range(4):
        bn = BatchNormAct2d(64, relu=True).cuda()
        xi = x.clone().requires_grad_(True)
        y = bn(xi)
        y.backward(dy)
        outs.append((y.detach().clone(), xi.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone(), bn.running_var.clone()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)
