"""Fused operand staging (gather / im2col + cast, cuda/operand_stage_sm100.cu), strided + split-K tcgen05 GEMM and the
tile-mode convolution built from them: each against a plain PyTorch fp32 reference of the same op."""
import pytest
import torch
import torch.nn.functional as F

from cpd_b200 import Quant_Conv, Quant_Linear, float_quantize, quant_gemm
from cpd_b200.ops import native

pytestmark = pytest.mark.gpu


def _cols_ref(x, kh, kw, stride, pad, dil):
    """F.unfold reordered to this repo's tap order (kh, kw, c): [B*OH*OW, KH*KW*C]."""
    B, C, H, W = x.shape
    cols = F.unfold(x.contiguous(), (kh, kw), dilation=dil, padding=pad, stride=stride)        # [B, C*kh*kw, L]
    L = cols.shape[-1]
    return cols.view(B, C, kh * kw, L).permute(0, 3, 2, 1).reshape(B * L, kh * kw * C)


@pytest.mark.parametrize("om,dt", [(native.OPERAND_BF16, torch.bfloat16), (native.OPERAND_TF32, torch.float32),
                                   (native.OPERAND_E4M3, torch.float8_e4m3fn), (native.OPERAND_E5M2, torch.float8_e5m2)])
def test_stage_matrix_equals_torch_cast_for_any_strides(om, dt):
    torch.manual_seed(0)
    x = torch.randn(200, 77, device="cuda") * 3
    x[3, 5], x[7, 9] = 1e6, -1e6                                # fp8: must saturate, not become NaN / Inf
    for view in (x, x.t(), x[::2, 3:50], x.t()[5:70:3]):
        got = native.stage_matrix(view, om)
        R, K = view.shape
        assert got.shape[0] == R and got.shape[1] % 16 == 0 and got.shape[1] >= K
        want = view.contiguous()
        if dt in (torch.float8_e4m3fn, torch.float8_e5m2):
            lim = 448.0 if dt == torch.float8_e4m3fn else 57344.0
            want = want.clamp(-lim, lim)
        assert torch.equal(got[:, :K].float(), want.to(dt).float())
        assert bool((got[:, K:].float() == 0).all())


@pytest.mark.parametrize("chan_last", [False, True])
@pytest.mark.parametrize("cfg", [(3, 1, 1, 1), (3, 2, 1, 1), (1, 1, 0, 1), (1, 2, 0, 1), (3, 1, 2, 2), (5, 2, 2, 1)])
def test_stage_im2col_equals_unfold(chan_last, cfg):
    k, stride, pad, dil = cfg
    torch.manual_seed(1)
    x = torch.randn(3, 20, 13, 11, device="cuda")
    if chan_last:
        x = x.contiguous(memory_format=torch.channels_last)
    OH = (13 + 2 * pad - dil * (k - 1) - 1) // stride + 1
    OW = (11 + 2 * pad - dil * (k - 1) - 1) // stride + 1
    want = _cols_ref(x, k, k, stride, pad, dil)
    K = k * k * 20
    got = native.stage_im2col(x, k, k, (OH, OW), stride, pad, dil, native.OPERAND_TF32)
    assert torch.equal(got[:, :K], want)
    got_t = native.stage_im2col(x, k, k, (OH, OW), stride, pad, dil, native.OPERAND_TF32, taps_are_rows=True)
    assert torch.equal(got_t[:, :want.shape[0]], want.t())
    got_b = native.stage_im2col(x, k, k, (OH, OW), stride, pad, dil, native.OPERAND_BF16)
    assert torch.equal(got_b[:, :K].float(), want.bfloat16().float())


def test_strided_padded_operands_and_split_k_match_matmul():
    torch.manual_seed(2)
    M, N, K = 200, 96, 1000                                        # K % 16 != 0: the TMA descriptor zero-fills the tail
    a, b = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
    ref = a.bfloat16().float() @ b.bfloat16().float()
    c = quant_gemm(a, b, man=23, exp=8, mode="tile", operands="bf16")
    assert (c - ref).abs().max() <= 2e-5 * ref.abs().max() * 8
    # split-K (ordered fold of the partial matrices) on a skinny output with a long reduction
    Kl = 16384
    a, b = torch.randn(64, Kl, device="cuda"), torch.randn(Kl, 72, device="cuda")
    ref = a.bfloat16().float().double() @ b.bfloat16().float().double()
    ao, bo = native.stage_matrix(a, native.OPERAND_BF16), native.stage_matrix(b.t(), native.OPERAND_BF16)
    for splits in (1, 4, 16):
        c = native.gemm_tc(ao, bo, 8, 23, False, 128, native.OPERAND_BF16, None, k=Kl, splits=splits)
        assert (c.double() - ref).abs().max() <= 1e-4 * ref.abs().max(), splits
    # with a narrow accumulator the split fold is the ordered quantised fold: every output is a member of the format
    c = native.gemm_tc(ao, bo, 5, 10, False, 128, native.OPERAND_BF16, None, k=Kl, splits=8)
    assert torch.equal(float_quantize(c, 5, 10, "ieee"), c)
    assert (c.double() - ref).abs().max() <= 0.02 * ref.abs().max()


@pytest.mark.parametrize("e,m,kahan", [(5, 10, False), (4, 3, False), (5, 2, True), (6, 9, False)])
@pytest.mark.parametrize("splits", [3, 8])
def test_split_k_is_the_ordered_quantised_fold_bit_for_bit(e, m, kahan, splits):
    """Exact-sum operands (multiples of 1/8): the kernel's split-K result must equal the C++ CPU model of the same
    partition (whole K-slices per split, splits folded in order with a cast after every add) bit for bit."""
    torch.manual_seed(splits + e)
    M, N, K = 96, 200, 4096
    a = (torch.randint(-4, 5, (M, K), device="cuda").float() / 8)
    b = (torch.randint(-4, 5, (N, K), device="cuda").float() / 8)
    ao, bo = native.stage_matrix(a, native.OPERAND_BF16), native.stage_matrix(b, native.OPERAND_BF16)
    got = native.gemm_tc(ao, bo, e, m, kahan, 128, native.OPERAND_BF16, None, k=K, splits=splits, variant=native.GEMM_1CTA).cpu()
    want = native.gemm_tile_cpu(a.cpu(), b.cpu().t(), e, m, rounding="ieee", kahan=kahan, k_slice=128, splits=splits)
    assert bool(((got == want) | (got.isnan() & want.isnan())).all()), float((got != want).float().mean())


@pytest.mark.parametrize("chan_last", [False, True])
@pytest.mark.parametrize("cfg", [(16, 32, 3, 1, 1), (32, 64, 3, 2, 1), (16, 24, 1, 2, 0), (3, 64, 3, 1, 1)])
def test_tile_conv_forward_backward_match_conv2d(chan_last, cfg):
    cin, cout, k, stride, pad = cfg
    torch.manual_seed(3)
    conv = Quant_Conv(cin, cout, k, stride=stride, padding=pad, bias=True, exp=8, man=23, mode="tile", operands="bf16").cuda()
    x = torch.randn(8, cin, 16, 16, device="cuda")
    if chan_last:
        x = x.contiguous(memory_format=torch.channels_last)
        conv = conv.to(memory_format=torch.channels_last)
    x.requires_grad_(True)
    y = conv(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    xb = x.detach().bfloat16().float().requires_grad_(True)
    wb = conv.weight.detach().bfloat16().float().requires_grad_(True)
    bb = conv.bias.detach().clone().requires_grad_(True)
    yr = F.conv2d(xb, wb, bb, stride=stride, padding=pad)
    assert y.shape == yr.shape
    assert (y - yr).abs().max() <= 1e-4 * yr.abs().max()
    # backward reference on the bf16-rounded gradient (the kernel casts dy as an operand too)
    dyb = dy.bfloat16().float()
    dxr, dwr = torch.autograd.grad(yr, (xb, wb), dyb, retain_graph=True)
    assert (x.grad - dxr).abs().max() <= 2e-4 * dxr.abs().max()
    assert (conv.weight.grad - dwr).abs().max() <= 2e-4 * dwr.abs().max()
    dbr = dy.sum((0, 2, 3))
    assert (conv.bias.grad - dbr).abs().max() <= 1e-4 * dbr.abs().max()


def test_tile_linear_forward_backward_match_matmul():
    torch.manual_seed(4)
    lin = Quant_Linear(300, 70, bias=True, exp=8, man=23, mode="tile").cuda()
    x = torch.randn(500, 300, device="cuda", requires_grad=True)
    y = lin(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    # (8,23): tf32 operands by default
    ref = F.linear(x.detach(), lin.weight.detach(), lin.bias.detach())
    assert (y - ref).abs().max() <= 3e-3 * ref.abs().max()
    assert (x.grad - dy @ lin.weight.detach()).abs().max() <= 3e-3 * (dy @ lin.weight.detach()).abs().max()
    assert (lin.weight.grad - dy.t() @ x.detach()).abs().max() <= 3e-3 * (dy.t() @ x.detach()).abs().max()
    assert (lin.bias.grad - dy.sum(0)).abs().max() <= 1e-4 * dy.sum(0).abs().max()
