"""APS + ordered/Kahan fold and LARS (C++ CPU ops) vs python oracles — SURVEY §4 row 3."""
import math

import numpy as np
import pytest
import torch

from cpd_b200.ops import native
from cpd_b200.parallel.arena import ArenaLayout
from oracle import aps_fold_oracle


def _make(W, numels, seed, scale_pows=(-20, 4)):
    rng = np.random.default_rng(seed)
    layout = ArenaLayout(numels)
    arenas = []
    for r in range(W):
        flat = torch.zeros(layout.total)
        for t, n in enumerate(numels):
            sc = 2.0 ** rng.integers(*scale_pows)
            flat[layout.offsets[t]:layout.offsets[t] + n] = torch.from_numpy((rng.standard_normal(n) * sc).astype(np.float32))
        arenas.append(flat)
    return layout, arenas


@pytest.mark.parametrize("W", [1, 2, 8])
@pytest.mark.parametrize("e,m", [(4, 3), (5, 2), (5, 10)])
@pytest.mark.parametrize("aps,kahan", [(True, False), (True, True), (False, False), (False, True)])
def test_fold_matches_oracle(W, e, m, aps, kahan):
    numels = [10, 64, 1, 4097, 300]
    layout, arenas = _make(W, numels, seed=W * 100 + e)
    tab = layout.table("cpu")
    out = torch.empty(layout.total)
    shifts = torch.zeros(len(numels), dtype=torch.int32)
    native.aps_fold(arenas, out, tab.seg_off, e, m, aps, kahan, float(W), shifts_out=shifts)
    for t, n in enumerate(numels):
        sl = slice(layout.offsets[t], layout.offsets[t] + n)
        want, sh = aps_fold_oracle([a[sl].numpy() for a in arenas], e, m, aps, kahan, float(W))
        assert int(shifts[t]) == sh
        got = out[sl].numpy()
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        assert same.all(), (t, W, e, m, aps, kahan)
    # padding stays zero
    assert float(out[layout.offsets[0] + numels[0]:layout.offsets[1]].abs().sum()) == 0.0


def test_fold_all_zero_tensor_and_extremes_are_guarded():
    layout = ArenaLayout([32, 32, 32])
    a = torch.zeros(layout.total)
    a[layout.offsets[1]:layout.offsets[1] + 32] = 1e-38          # tiny: shift clamps
    a[layout.offsets[2]:layout.offsets[2] + 32] = 1e30           # huge: negative shift
    out = torch.empty_like(a)
    sh = torch.zeros(3, dtype=torch.int32)
    native.aps_fold([a, a.clone()], out, layout.table("cpu").seg_off, 5, 2, True, False, 2.0, shifts_out=sh)
    assert sh[0].item() == 0 and torch.isfinite(out).all()       # the reference would produce NaN here (0 * inf)
    assert (out[:32] == 0).all()
    assert sh[2].item() < 0 and torch.allclose(out[layout.offsets[2]:layout.offsets[2] + 32], torch.full((32,), 2e30), rtol=0.2)


def test_aps_prevents_underflow_that_plain_low_precision_suffers():
    """The point of APS: small gradients survive an (5,2) sum only when shifted."""
    layout = ArenaLayout([4096])
    g = [torch.full((4096,), 1e-6) * (1 + 0.1 * r) for r in range(8)]
    tab = layout.table("cpu")
    exact = sum(x.double() for x in g).float()
    with_aps, without = torch.empty(4096), torch.empty(4096)
    native.aps_fold(g, with_aps, tab.seg_off, 5, 2, True, False, 8.0)
    native.aps_fold(g, without, tab.seg_off, 5, 2, False, False, 8.0)
    err_aps = ((with_aps - exact).abs() / exact).max()
    err_plain = ((without - exact).abs() / exact).max()
    assert err_aps < 0.15 and err_plain > 0.5


def test_rounding_mode_reaches_the_gradient_fold():
    """PrecisionConfig.rounding must act on the gradient path.  Quirk B (truncating subnormal shift) makes the
    reference cast of 0.5*ulp_min*(1+2^-20) zero where IEEE rounds up; the fold of one source shows which cast ran."""
    from cpd_b200 import PrecisionConfig, float_quantize
    from cpd_b200.parallel import GradReducer, ModelArena
    layout = ArenaLayout([64])
    ulp_min = 2.0 ** (1 - 7 - 3)                          # (4,3): smallest subnormal
    x = torch.full((layout.total,), 0.5 * ulp_min * (1 + 2.0 ** -20))
    tab = layout.table("cpu")
    ref, ieee = torch.empty(layout.total), torch.empty(layout.total)
    native.aps_fold([x], ref, tab.seg_off, 4, 3, False, False, 1.0, rounding="reference")
    native.aps_fold([x], ieee, tab.seg_off, 4, 3, False, False, 1.0, rounding="ieee")
    assert torch.equal(ref, float_quantize(x, 4, 3, "reference")) and float(ref[0]) == 0.0
    assert torch.equal(ieee, float_quantize(x, 4, 3, "ieee")) and float(ieee[0]) == ulp_min
    # ... and through the public reducer (world size 1, no APS so the raw value reaches the cast)
    lin = torch.nn.Linear(8, 8, bias=False)
    arena = ModelArena(list(lin.parameters()))
    for name, want in (("reference", 0.0), ("ieee", ulp_min)):
        arena.grad_flat.fill_(float(x[0]))
        GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, rounding=name)).reduce()
        assert float(arena.grad_flat[0]) == want, name


def test_fp32_identity_is_plain_sum():
    layout = ArenaLayout([1000])
    g = [torch.randn(layout.total) for _ in range(4)]
    out = torch.empty(layout.total)
    native.aps_fold(g, out, layout.table("cpu").seg_off, 8, 23, False, False, 4.0)
    assert torch.equal(out, ((g[0] + g[1]) + g[2]) + g[3])


def _lars_oracle(w, g, buf, lr, mom, wd, coeff):
    wn, gn = w.double().norm().item(), g.double().norm().item()
    local = coeff * wn / (gn + wd * wn)
    buf = mom * buf + lr * local * (g + wd * w)
    return w - buf, buf


def test_lars_step_matches_reference_formula():
    numels = [64, 3000, 10, 5000]
    layout = ArenaLayout(numels)
    torch.manual_seed(0)
    master = torch.zeros(layout.total)
    grad = torch.zeros(layout.total)
    for t, n in enumerate(numels):
        master[layout.offsets[t]:layout.offsets[t] + n] = torch.randn(n)
        grad[layout.offsets[t]:layout.offsets[t] + n] = torch.randn(n) * 0.01
    mom = torch.zeros(layout.total)
    model = torch.zeros(layout.total)
    m0, g0 = master.clone(), grad.clone()
    tab = layout.table("cpu")
    bufs = [torch.zeros(n) for n in numels]
    ws = [m0[layout.offsets[t]:layout.offsets[t] + n].clone() for t, n in enumerate(numels)]
    for step in range(3):
        native.lars_step(master, model, grad, mom, tab.seg_off, lr=0.1, momentum=0.9, wd=1e-4, coeff=0.001, use_lars=True)
        for t, n in enumerate(numels):
            ws[t], bufs[t] = _lars_oracle(ws[t], g0[layout.offsets[t]:layout.offsets[t] + n], bufs[t], 0.1, 0.9, 1e-4, 0.001)
    for t, n in enumerate(numels):
        sl = slice(layout.offsets[t], layout.offsets[t] + n)
        assert torch.allclose(master[sl], ws[t], rtol=1e-5, atol=1e-7)
        assert torch.equal(model[sl], master[sl])


@pytest.mark.parametrize("nesterov", [False, True])
def test_sgd_mode_matches_torch_optim(nesterov):
    numels = [100, 257]
    layout = ArenaLayout(numels)
    torch.manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(n)) for n in numels]
    opt = torch.optim.SGD(ps, lr=0.05, momentum=0.9, weight_decay=5e-4, nesterov=nesterov)
    master = torch.zeros(layout.total)
    for t, p in enumerate(ps):
        master[layout.offsets[t]:layout.offsets[t] + numels[t]] = p.detach()
    mom = torch.zeros(layout.total)
    tab = layout.table("cpu")
    for step in range(4):
        grad = torch.zeros(layout.total)
        for t, p in enumerate(ps):
            p.grad = torch.randn(numels[t])
            grad[layout.offsets[t]:layout.offsets[t] + numels[t]] = p.grad
        opt.step()
        native.lars_step(master, None, grad, mom, tab.seg_off, lr=0.05, momentum=0.9, wd=5e-4, use_lars=False, nesterov=nesterov)
    for t, p in enumerate(ps):
        assert torch.allclose(master[layout.offsets[t]:layout.offsets[t] + numels[t]], p.detach(), rtol=1e-5, atol=1e-6)
