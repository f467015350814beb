"""The tile-mode convolution / linear autograd functions (quant/modules.py) on the CPU, with the native staging and
tensor-core entry points replaced by a numpy model that mirrors operand_stage_sm100.cu's address generators line by
line.  This pins the DESCRIPTOR logic — tap order (kh, kw, c), the flipped filter through negative strides and a base
offset, the transposed-convolution padding, input dilation for strided layers, dy^T / transposed im2col for the weight
gradient, channels_last vs NCHW strides — against torch's own conv2d forward and backward, without a GPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cpd_b200.ops import native
from cpd_b200.quant import Quant_Conv, Quant_Linear


def _flat(src):
    """The tensor's whole storage as a 1-D numpy array + the element offset of src[0,...,0] in it."""
    st = torch.empty(0, dtype=src.dtype).set_(src.untyped_storage())
    return st.numpy(), src.storage_offset()


class _Spec:
    def __init__(self, mat, K):
        self.mat, self.K = mat, K


def _gather(src, R, sizes, strides, s_row, om, base=0, rows_fast=None):
    flat, off = _flat(src)
    n1, n2, n3 = sizes
    s1, s2, s3 = strides
    r = np.arange(R, dtype=np.int64)[:, None]
    k = np.arange(n1 * n2 * n3, dtype=np.int64)[None, :]
    n23 = n2 * n3                                            # stage_fetch(), mode 0
    k1, rem = k // n23, k % n23
    k2, k3 = rem // n3, rem % n3
    idx = off + base + r * s_row + k1 * s1 + k2 * s2 + k3 * s3
    assert idx.min() >= 0 and idx.max() < flat.size
    return _Spec(torch.from_numpy(flat[idx].astype(np.float32)), n1 * n2 * n3)


def _matrix(x, om):
    return _gather(x, x.shape[0], (x.shape[1], 1, 1), (x.stride(1), 0, 0), x.stride(0), om)


def _im2col(x, kh, kw, out_hw, stride, pad, dil, om, in_dil=1, taps_are_rows=False):
    flat, off = _flat(x)
    B, C, H, W = x.shape
    OH, OW = out_hw
    sb, sc, sh, sw = x.stride()
    pixels, taps = B * OH * OW, kh * kw * C
    pix = np.arange(pixels, dtype=np.int64)[:, None]
    tap = np.arange(taps, dtype=np.int64)[None, :]
    ohw = OH * OW                                            # stage_fetch(), mode 1
    b, prem = pix // ohw, pix % ohw
    oh, ow = prem // OW, prem % OW
    kwc = kw * C
    khi, trem = tap // kwc, tap % kwc
    kwi, c = trem // C, trem % C
    ih = oh * stride - pad + khi * dil
    iw = ow * stride - pad + kwi * dil
    ok = np.ones_like(ih, dtype=bool)
    if in_dil > 1:
        ok &= (ih >= 0) & (iw >= 0) & (ih % in_dil == 0) & (iw % in_dil == 0)
        ih, iw = ih // in_dil, iw // in_dil
    ok &= (ih >= 0) & (ih < H) & (iw >= 0) & (iw < W)
    idx = off + b * sb + np.clip(ih, 0, H - 1) * sh + np.clip(iw, 0, W - 1) * sw + c * sc
    mat = np.where(ok, flat[idx], 0.0).astype(np.float32)
    if taps_are_rows:
        mat = mat.T.copy()
    return _Spec(torch.from_numpy(mat), mat.shape[1])


def _stage(*specs):
    return tuple(s.mat for s in specs) if len(specs) > 1 else specs[0].mat


def _gemm_tc(a, bt, exp, man, kahan=False, k_slice=128, operand_mode=0, bias=None, out=None, variant=0, k=None, splits=1):
    k = a.shape[1] if k is None else k
    c = a[:, :k].double() @ bt[:, :k].double().t()
    if bias is not None:
        c = c + bias.detach().double()
    return c.float()


@pytest.fixture
def cpu_model(monkeypatch):
    from cpd_b200.quant import functional as qfn
    monkeypatch.setattr(qfn, "_CPU_MODEL", False)            # drive the GPU path's descriptor logic, not the CPU model
    monkeypatch.setattr(native, "spec_gather", _gather)
    monkeypatch.setattr(native, "spec_matrix", _matrix)
    monkeypatch.setattr(native, "spec_im2col", _im2col)
    monkeypatch.setattr(native, "stage", _stage)
    monkeypatch.setattr(native, "gemm_tc", _gemm_tc)
    monkeypatch.setattr(native, "colsum_quantize", lambda x2d, e, m, rounding=0: x2d.sum(0))


@pytest.mark.parametrize("chan_last", [False, True])
@pytest.mark.parametrize("cfg", [(5, 7, 3, 1, 1, 1), (4, 6, 3, 2, 1, 1), (6, 5, 1, 2, 0, 1), (3, 8, 3, 1, 2, 2), (4, 4, 5, 2, 2, 1),
                                 (2, 3, 3, 3, 1, 1)])
def test_tile_conv_descriptors_reproduce_conv2d_forward_and_backward(cpu_model, chan_last, cfg):
    cin, cout, k, stride, pad, dil = cfg
    torch.manual_seed(cin * 10 + k)
    conv = Quant_Conv(cin, cout, k, stride=stride, padding=pad, dilation=dil, bias=True, exp=8, man=23, mode="tile", operands="tf32")
    x = torch.randn(2, cin, 11, 9)
    if chan_last:
        x = x.contiguous(memory_format=torch.channels_last)
        conv = conv.to(memory_format=torch.channels_last)
    x.requires_grad_(True)
    y = conv(x)
    xr = x.detach().clone().requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    br = conv.bias.detach().clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, stride=stride, padding=pad, dilation=dil)
    assert y.shape == yr.shape
    assert torch.allclose(y, yr, rtol=1e-5, atol=1e-5)
    for dy in (torch.randn_like(yr), torch.randn_like(yr).contiguous(memory_format=torch.channels_last),
               torch.ones(1).expand_as(yr)):                 # the last one: the stride-0 gradient of y.sum()
        for t in (x, conv.weight, conv.bias, xr, wr, br):
            t.grad = None
        y.backward(dy, retain_graph=True)
        yr.backward(dy, retain_graph=True)
        assert torch.allclose(x.grad, xr.grad, rtol=1e-4, atol=1e-4)
        assert torch.allclose(conv.weight.grad, wr.grad, rtol=1e-4, atol=1e-4)
        assert torch.allclose(conv.bias.grad, br.grad, rtol=1e-4, atol=1e-4)


def test_tile_linear_descriptors_reproduce_matmul(cpu_model):
    torch.manual_seed(0)
    lin = Quant_Linear(13, 7, bias=True, exp=8, man=23, mode="tile")
    x = torch.randn(10, 13, requires_grad=True)
    y = lin(x)
    ref = F.linear(x.detach(), lin.weight.detach(), lin.bias.detach())
    assert torch.allclose(y, ref, rtol=1e-5, atol=1e-6)
    dy = torch.randn_like(y)
    y.backward(dy)
    assert torch.allclose(x.grad, dy @ lin.weight.detach(), rtol=1e-5, atol=1e-5)
    assert torch.allclose(lin.weight.grad, dy.t() @ x.detach(), rtol=1e-5, atol=1e-5)
    assert torch.allclose(lin.bias.grad, dy.sum(0), rtol=1e-5, atol=1e-5)


def _gemm_conv_model(x_nhwc, bt, kh, kw, out_hw, stride, pad, dil, exp, man, kahan=False, k_slice=128, operand_mode=0, bias=None,
                     variant=0, splits=1):
    """native.gemm_conv as a torch convolution: x [B,H,W,C], bt [O, (kh,kw,c)] -> [B*OH*OW, O]."""
    B, H, W, C = x_nhwc.shape
    O = bt.shape[0]
    w = bt[:, :kh * kw * C].reshape(O, kh, kw, C).permute(0, 3, 1, 2)
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2).double(), w.double(), None if bias is None else bias.detach().double(), stride=stride,
                 padding=pad, dilation=dil)
    assert tuple(y.shape[-2:]) == tuple(out_hw)
    return y.permute(0, 2, 3, 1).reshape(-1, O).float()


def _gemm_wgrad_model(dy_rows, x_nhwc, kh, kw, out_hw, stride, pad, dil, exp, man, kahan=False, k_slice=128, splits=1, _lbo=0, _sbo=0):
    """native.gemm_wgrad: dW [O, (kh,kw,c)] = sum_pixels dy[pixel, o] * x-tap[pixel, (kh,kw,c)]."""
    B, H, W, C = x_nhwc.shape
    O = dy_rows.shape[1]
    cols = F.unfold(x_nhwc.permute(0, 3, 1, 2).double(), (kh, kw), dilation=dil, padding=pad, stride=stride)       # [B, C*kh*kw, L]
    L = cols.shape[-1]
    cols = cols.view(B, C, kh * kw, L).permute(0, 3, 2, 1).reshape(B * L, kh * kw * C)
    assert dy_rows.shape[0] == B * L and dy_rows.is_contiguous() and x_nhwc.is_contiguous()
    return (dy_rows.double().t() @ cols).float()


@pytest.mark.parametrize("dy_kind", ["like", "nchw", "expanded"])
@pytest.mark.parametrize("cfg", [(8, 12, 3, 1, 1, 1), (8, 8, 3, 2, 1, 1), (8, 16, 1, 2, 0, 1), (8, 8, 3, 1, 2, 2), (8, 12, 5, 1, 2, 1)])
def test_implicit_conv_glue_reproduces_conv2d_forward_and_backward(cpu_model, monkeypatch, cfg, dy_kind):
    """The IMPLICIT-GEMM branch of Quant_ConvTileFunction (cast-only staging, the bf16 x kept for backward, dy cast shared by
    the data- and weight-gradient GEMMs, staged fallback for the strided data gradient, a non-channels_last dy) with
    native.gemm_conv / gemm_wgrad replaced by torch models of what the kernels compute."""
    cin, cout, k, stride, pad, dil = cfg
    calls = {"conv": 0, "wgrad": 0}

    def conv_model(*a, **kw):
        calls["conv"] += 1
        return _gemm_conv_model(*a, **kw)

    def wgrad_model(*a, **kw):
        calls["wgrad"] += 1
        return _gemm_wgrad_model(*a, **kw)

    monkeypatch.setattr(native, "conv_implicit_ok", lambda *a, **kw: True)
    monkeypatch.setattr(native, "gemm_conv", conv_model)
    monkeypatch.setattr(native, "gemm_wgrad", wgrad_model)
    torch.manual_seed(cin + k)
    conv = Quant_Conv(cin, cout, k, stride=stride, padding=pad, dilation=dil, bias=True, exp=8, man=23, mode="tile", operands="bf16")
    conv = conv.to(memory_format=torch.channels_last)
    x = torch.randn(3, cin, 10, 9).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = conv(x)
    xr = x.detach().clone().requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    br = conv.bias.detach().clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, stride=stride, padding=pad, dilation=dil)
    assert torch.allclose(y, yr, rtol=1e-5, atol=1e-5) and calls["conv"] == 1
    dy = {"like": torch.randn_like(y), "nchw": torch.randn(yr.shape), "expanded": torch.ones(1).expand_as(yr)}[dy_kind]
    y.backward(dy)
    yr.backward(dy)
    assert torch.allclose(x.grad, xr.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(conv.weight.grad, wr.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(conv.bias.grad, br.grad, rtol=1e-4, atol=1e-4)
    assert calls["wgrad"] == (1 if cout % 8 == 0 else 0)               # O * 2 bytes must be a multiple of 16 for dy's tensor map
    # data gradient: implicit for stride 1 when dy arrives (or is made, for the implicit weight gradient) channels_last;
    # strided layers take the staged zero-upsampled path
    assert calls["conv"] == (2 if stride == 1 and (dy_kind == "like" or cout % 8 == 0) else 1)
