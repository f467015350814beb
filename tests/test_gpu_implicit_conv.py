"""Implicit-GEMM convolution (im2col-mode TMA A operand of the tcgen05 GEMMs): against the explicit im2col matrix fed
to the same kernel (bit for bit: same operand values, same K order, same K-slices) and against F.conv2d in fp32."""
import pytest
import torch
import torch.nn.functional as F

from cpd_b200 import Quant_Conv
from cpd_b200.ops import native
from cpd_b200.quant import modules as qmod

pytestmark = pytest.mark.gpu

# (B, C, H, W, O, k, stride, pad, dil)
GEOMS = [(4, 64, 16, 16, 64, 3, 1, 1, 1),          # M = 1024: whole tiles
         (3, 64, 10, 14, 96, 3, 1, 1, 1),          # M = 420: row / image wraps inside a tile, M tail, N tail
         (4, 128, 16, 16, 128, 3, 2, 1, 1),        # strided
         (5, 64, 9, 9, 32, 1, 2, 0, 1),            # 1x1 stride-2 shortcut, odd extent
         (2, 64, 12, 12, 64, 3, 1, 2, 2),          # dilated
         (2, 64, 8, 8, 64, 5, 1, 2, 1),            # 5x5
         (2, 192, 7, 5, 40, 3, 1, 0, 1)]           # no padding, three channel blocks per tap


def _operands(g, dt):
    B, C, H, W, O, k, stride, pad, dil = g
    torch.manual_seed(11)
    x = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    w = torch.randn(O, C, k, k, device="cuda") / (C * k * k) ** 0.5
    OH = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    OW = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(dt)
    wt = w.permute(0, 2, 3, 1).reshape(O, k * k * C).contiguous().to(dt)                  # (kh, kw, c)
    return x, w, x_nhwc, wt, (OH, OW)


@pytest.mark.parametrize("variant", [native.GEMM_1CTA, native.GEMM_2CTA])
@pytest.mark.parametrize("g", GEOMS)
def test_implicit_gemm_equals_explicit_im2col_gemm_bitwise(g, variant):
    B, C, H, W, O, k, stride, pad, dil = g
    x, w, x_nhwc, wt, (OH, OW) = _operands(g, torch.bfloat16)
    bias = torch.randn(O, device="cuda")
    cols = native.stage_im2col(x, k, k, (OH, OW), stride, pad, dil, native.OPERAND_BF16)            # explicit matrix, same tap order
    for (e, m, ks) in ((8, 23, 0), (5, 10, 128)):
        if variant == native.GEMM_2CTA and (e, m) == (8, 23):
            continue                                                                                 # packed accumulators only
        ref = native.gemm_tc(cols, wt, e, m, False, ks, native.OPERAND_BF16, bias, variant=variant, k=k * k * C)
        got = native.gemm_conv(x_nhwc, wt, k, k, (OH, OW), stride, pad, dil, e, m, False, ks, native.OPERAND_BF16, bias, variant=variant)
        assert got.shape == ref.shape == (B * OH * OW, O)
        assert torch.equal(got, ref), f"{(got != ref).sum().item()} of {got.numel()} differ, max {(got - ref).abs().max().item()}"
    # and the explicit matrix is the right one: fp32 convolution of the bf16-rounded operands
    yr = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), bias, stride=stride, padding=pad, dilation=dil)
    got = native.gemm_conv(x_nhwc, wt, k, k, (OH, OW), stride, pad, dil, 8, 23, False, 0, native.OPERAND_BF16, bias, variant=native.GEMM_1CTA)
    got = got.view(B, OH, OW, O).permute(0, 3, 1, 2)
    assert (got - yr).abs().max() <= 1e-4 * yr.abs().max()


def test_implicit_gemm_split_k_equals_explicit_split_k():
    g = (2, 256, 8, 8, 128, 3, 1, 1, 1)                      # 128 output positions, K = 2304: the under-filled deep-layer shape
    B, C, H, W, O, k, stride, pad, dil = g
    x, w, x_nhwc, wt, (OH, OW) = _operands(g, torch.bfloat16)
    cols = native.stage_im2col(x, k, k, (OH, OW), stride, pad, dil, native.OPERAND_BF16)
    for splits in (2, 4):
        ref = native.gemm_tc(cols, wt, 5, 10, False, 128, native.OPERAND_BF16, None, k=k * k * C, splits=splits)
        got = native.gemm_conv(x_nhwc, wt, k, k, (OH, OW), stride, pad, dil, 5, 10, False, 128, native.OPERAND_BF16, splits=splits)
        assert torch.equal(got, ref)


def test_implicit_gemm_fp8_operands():
    g = (2, 128, 12, 12, 64, 3, 1, 1, 1)
    B, C, H, W, O, k, stride, pad, dil = g
    x, w, x_nhwc, wt, (OH, OW) = _operands(g, torch.float8_e4m3fn)
    cols = native.stage_im2col(x, k, k, (OH, OW), stride, pad, dil, native.OPERAND_E4M3)
    ref = native.gemm_tc(cols, wt, 5, 10, False, 256, native.OPERAND_E4M3, None, k=k * k * C)
    got = native.gemm_conv(x_nhwc, wt, k, k, (OH, OW), stride, pad, dil, 5, 10, False, 256, native.OPERAND_E4M3)
    assert torch.equal(got, ref)


def test_implicit_gemm_rejects_what_it_does_not_cover():
    assert not native.conv_implicit_ok(3, 3, 3, 1, 1, 1, native.OPERAND_BF16)          # stem: 3 channels
    assert not native.conv_implicit_ok(64, 3, 3, 1, 1, 1, native.OPERAND_E4M3)         # fp8 needs 128-channel blocks
    assert native.conv_implicit_ok(128, 3, 3, 2, 1, 1, native.OPERAND_E4M3)
    x = torch.zeros(1, 4, 4, 48, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        native.gemm_conv(x, torch.zeros(8, 9 * 48, device="cuda", dtype=torch.bfloat16), 3, 3, (4, 4), 1, 1, 1, 5, 10)


@pytest.mark.parametrize("cfg", [(64, 64, 3, 1, 1), (64, 128, 3, 2, 1), (64, 128, 1, 2, 0), (128, 64, 3, 1, 1)])
def test_quant_conv_implicit_equals_staged_path(cfg, monkeypatch):
    """Quant_Conv with the implicit path on == off, bit for bit (forward, data gradient, weight gradient)."""
    cin, cout, k, stride, pad = cfg
    outs = []
    for implicit in (False, True):
        monkeypatch.setattr(qmod, "_IMPLICIT", implicit)
        torch.manual_seed(5)
        conv = Quant_Conv(cin, cout, k, stride=stride, padding=pad, bias=True, exp=5, man=10, mode="tile", operands="bf16").cuda()
        conv = conv.to(memory_format=torch.channels_last)
        x = torch.randn(6, cin, 12, 12, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
        before = native._C.launch_count()
        y = conv(x)
        fwd_launches = native._C.launch_count() - before
        dy = torch.randn(y.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(9)).contiguous(memory_format=torch.channels_last)
        y.backward(dy)
        outs.append((y.detach().clone(), x.grad.clone(), conv.weight.grad.clone(), fwd_launches))
    for a, b in zip(outs[0][:3], outs[1][:3]):
        assert torch.equal(a, b)
    from cpd_b200.quant.functional import splits_for
    oh = (12 + 2 * pad - (k - 1) - 1) // stride + 1
    assert outs[1][3] == 2 + (splits_for(6 * oh * oh, cout, k * k * cin, 128) > 1)      # cast + GEMM (+ split-K fold)


def _wgrad_operands(g):
    B, C, H, W, O, k, stride, pad, dil = g
    x, w, x_nhwc, wt, (OH, OW) = _operands(g, torch.bfloat16)
    dy = torch.randn(B, O, OH, OW, device="cuda", generator=torch.Generator("cuda").manual_seed(3)).contiguous(memory_format=torch.channels_last)
    dy_rows = dy.permute(0, 2, 3, 1).reshape(B * OH * OW, O).contiguous().to(torch.bfloat16)
    return x, x_nhwc, dy, dy_rows, (OH, OW)


@pytest.mark.parametrize("g", [g for g in GEOMS if g[4] % 8 == 0])
def test_implicit_wgrad_equals_staged_wgrad_bitwise(g):
    """MN-major implicit weight gradient == the K-major GEMM on the explicitly transposed operands (same K order, same
    K-slices, same split-K fold) and == autograd's weight gradient in fp32."""
    B, C, H, W, O, k, stride, pad, dil = g
    x, x_nhwc, dy, dy_rows, (OH, OW) = _wgrad_operands(g)
    ds = dy.stride()
    a, bt = native.stage(native.spec_gather(dy, O, (B, OH, OW), (ds[0], ds[2], ds[3]), ds[1], native.OPERAND_BF16),
                         native.spec_im2col(x, k, k, (OH, OW), stride, pad, dil, native.OPERAND_BF16, taps_are_rows=True))
    Mred = B * OH * OW
    for (e, m, ks, splits) in ((8, 23, 0, 1), (5, 10, 128, 1), (5, 10, 128, 3)):
        ref = native.gemm_tc(a, bt, e, m, False, ks, native.OPERAND_BF16, None, variant=native.GEMM_1CTA, k=Mred, splits=splits)
        got = native.gemm_wgrad(dy_rows, x_nhwc, k, k, (OH, OW), stride, pad, dil, e, m, False, ks, splits=splits)
        assert got.shape == ref.shape == (O, k * k * C)
        assert torch.equal(got, ref), f"{(got != ref).sum().item()} of {got.numel()} differ, max {(got - ref).abs().max().item()}"
    wr = torch.zeros(O, C, k, k, device="cuda", requires_grad=True)
    yr = F.conv2d(x.bfloat16().float(), wr, None, stride=stride, padding=pad, dilation=dil)
    (dwr,) = torch.autograd.grad(yr, wr, dy.bfloat16().float())
    got = native.gemm_wgrad(dy_rows, x_nhwc, k, k, (OH, OW), stride, pad, dil, 8, 23, False, 0)
    got = got.view(O, k, k, C).permute(0, 3, 1, 2)
    assert (got - dwr).abs().max() <= 2e-4 * dwr.abs().max()
