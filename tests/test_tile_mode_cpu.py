"""CPU model of the tensor-core "tile" mode (quant_gemm / Quant_Linear / Quant_Conv with mode="tile" on CPU tensors):
operands rounded to the operand format, wide partial sums per K-slice, accumulator cast after every slice, split-K as
an ordered fold - the semantics the tcgen05 kernels implement, readable and testable without a GPU."""
import pytest
import torch
import torch.nn.functional as F

from cpd_b200 import Quant_Conv, Quant_Linear, float_quantize, quant_gemm


def _bf(t):
    return t.bfloat16().float()


def test_tile_gemm_cpu_fp32_accumulator_is_the_bf16_product():
    torch.manual_seed(0)
    a, b = torch.randn(33, 300), torch.randn(300, 17)
    c = quant_gemm(a, b, man=23, exp=8, mode="tile", operands="bf16")
    ref = (_bf(a).double() @ _bf(b).double()).float()
    assert (c - ref).abs().max() <= 2e-6 * ref.abs().max()


def test_tile_gemm_cpu_casts_once_per_k_slice():
    """One slice (k_slice >= K): a single cast of the exact product.  Two slices: s = Q(Q(p0) + Q(p1))."""
    torch.manual_seed(1)
    a, b = torch.randn(8, 256), torch.randn(256, 5)
    exact = lambda x, y: (_bf(x).double() @ _bf(y).double()).float()         # noqa: E731
    one = quant_gemm(a, b, man=10, exp=5, mode="tile", operands="bf16", k_slice=256)
    assert torch.equal(one, float_quantize(exact(a, b), 5, 10, "ieee"))
    two = quant_gemm(a, b, man=10, exp=5, mode="tile", operands="bf16", k_slice=128)
    p0, p1 = exact(a[:, :128], b[:128]), exact(a[:, 128:], b[128:])
    q = lambda t: float_quantize(t, 5, 10, "ieee")                            # noqa: E731
    assert torch.equal(two, q(q(p0) + q(p1)))


def test_tile_linear_cpu_forward_backward():
    torch.manual_seed(2)
    lin = Quant_Linear(40, 24, exp=8, man=23, mode="tile")
    x = torch.randn(16, 40, requires_grad=True)
    y = lin(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr, wr = x.detach().clone().requires_grad_(True), lin.weight.detach().clone().requires_grad_(True)
    yr = F.linear(xr, wr, lin.bias.detach())
    yr.backward(dy)
    for got, want in ((y, yr), (x.grad, xr.grad), (lin.weight.grad, wr.grad)):
        assert (got - want).abs().max() <= 3e-3 * want.abs().max()             # tf32 operands (default above 7 mantissa bits)
    assert torch.allclose(lin.bias.grad, dy.sum(0), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("cfg", [(4, 6, 3, 1, 1, 8), (4, 6, 3, 2, 1, 9), (5, 3, 1, 2, 0, 8), (3, 4, 3, 2, 0, 9), (2, 3, 5, 2, 2, 10), (3, 4, 3, 1, 2, 9)])
def test_tile_conv_cpu_matches_conv2d_on_the_rounded_operands(cfg):
    cin, cout, k, stride, pad, hw = cfg
    torch.manual_seed(3)
    dil = 2 if pad == 2 and k == 3 else 1
    conv = Quant_Conv(cin, cout, k, stride=stride, padding=pad, dilation=dil, bias=True, exp=8, man=23, mode="tile", operands="bf16")
    x = torch.randn(2, cin, hw, hw, requires_grad=True)
    y = conv(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr, wr = _bf(x.detach()).requires_grad_(True), _bf(conv.weight.detach()).requires_grad_(True)
    yr = F.conv2d(xr, wr, conv.bias.detach(), stride=stride, padding=pad, dilation=dil)
    dxr, dwr = torch.autograd.grad(yr, (xr, wr), _bf(dy))
    assert y.shape == yr.shape
    assert (y - yr).abs().max() <= 1e-5 * yr.abs().max()
    assert (x.grad - dxr).abs().max() <= 1e-5 * dxr.abs().max()
    assert (conv.weight.grad - dwr).abs().max() <= 1e-5 * dwr.abs().max()
    assert torch.allclose(conv.bias.grad, dy.sum((0, 2, 3)), rtol=1e-5, atol=1e-5)


def test_tile_conv_cpu_low_precision_accumulator_stays_in_format():
    torch.manual_seed(4)
    conv = Quant_Conv(8, 8, 3, padding=1, bias=False, exp=4, man=3, mode="tile", operands="e4m3", k_slice=16)
    y = conv(torch.randn(1, 8, 6, 6))
    assert torch.equal(y, float_quantize(y, 4, 3, "ieee"))                  # every output is a member of (4,3)


@pytest.mark.parametrize("e5m2", [False, True])
def test_block_scaled_operand_model(e5m2):
    """mx_round: power-of-two scale per 32 K-elements chosen so that the block's maximum just fits the fp8 range; blocks of
    very different magnitude keep their relative precision (what a per-tensor fp8 cast cannot do)."""
    from cpd_b200.quant.functional import mx_round
    torch.manual_seed(5)
    x = torch.randn(6, 100) * torch.logspace(-6, 6, 100 // 25 * 25 + 0, base=10.0)[:100].roll(3)
    x[0, :32] = 0.0                                                   # an all-zero block
    y = mx_round(x, e5m2)
    assert y.shape == x.shape and torch.equal(y[0, :32], torch.zeros(32))
    rel = 2.0 ** -(2 if e5m2 else 3)                                   # half an ulp of the block maximum's binade, relative
    xb, yb = torch.nn.functional.pad(x, (0, 28)).view(6, 4, 32), torch.nn.functional.pad(y, (0, 28)).view(6, 4, 32)
    amax = xb.abs().amax(-1, keepdim=True)
    assert ((xb - yb).abs() <= rel * amax).all()                      # error bounded relative to the BLOCK maximum
    assert torch.equal(mx_round(y, e5m2), y)                          # idempotent
    max_code = 57344.0 if e5m2 else 448.0
    ratio = yb.abs().amax(-1) / amax.squeeze(-1).clamp_min(1e-30)     # the maximum survives (no saturation, no underflow)
    assert ((ratio > 0.9) | (amax.squeeze(-1) == 0)).all() and max_code > 0


def test_block_scaled_tile_gemm_cpu():
    torch.manual_seed(6)
    a, b = torch.randn(9, 96) * 100.0, torch.randn(96, 7) * 1e-3
    from cpd_b200.quant.functional import mx_round
    c = quant_gemm(a, b, man=23, exp=8, mode="tile", operands="mxe4m3")
    ref = (mx_round(a).double() @ mx_round(b.t()).t().double()).float()
    assert (c - ref).abs().max() <= 2e-6 * ref.abs().max()
    plain = a @ b
    assert (c - plain).abs().max() <= 0.08 * plain.abs().max()       # fp8 operands: a few percent
