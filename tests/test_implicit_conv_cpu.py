"""Host-side logic of the implicit-GEMM convolution path (no GPU): coverage predicate, channels_last row view, and the
im2col-mode TMA coordinate arithmetic of the kernels' producer warp against a plain unfold."""
import numpy as np
import torch
import torch.nn.functional as F

from cpd_b200.ops import native
from cpd_b200.quant.functional import splits_for
from cpd_b200.quant.modules import _nhwc_rows


def test_coverage_predicate():
    ok = native.conv_implicit_ok
    assert ok(64, 3, 3, 1, 1, 1, native.OPERAND_BF16) and ok(512, 1, 1, 2, 0, 1, native.OPERAND_BF16)
    assert not ok(3, 3, 3, 1, 1, 1, native.OPERAND_BF16)              # stem
    assert not ok(96, 3, 3, 1, 1, 1, native.OPERAND_BF16)             # not whole 128-byte channel blocks
    assert not ok(64, 3, 3, 1, 1, 1, native.OPERAND_E4M3) and ok(128, 3, 3, 1, 1, 1, native.OPERAND_E4M3)
    assert not ok(48, 3, 3, 1, 1, 1, native.OPERAND_TF32) and ok(32, 3, 3, 1, 1, 1, native.OPERAND_TF32)
    assert not ok(64, 3, 3, 9, 1, 1, native.OPERAND_BF16) and not ok(64, 3, 3, 1, 200, 1, native.OPERAND_BF16)


def test_nhwc_rows_is_a_view_of_channels_last_only():
    x = torch.randn(2, 8, 3, 5).contiguous(memory_format=torch.channels_last)
    r = _nhwc_rows(x)
    assert r.shape == (30, 8) and r.data_ptr() == x.data_ptr()
    assert torch.equal(r, x.permute(0, 2, 3, 1).reshape(30, 8))
    assert _nhwc_rows(torch.randn(2, 8, 3, 5)) is None
    assert _nhwc_rows(x[:, :, :, ::2]) is None


def test_deep_layers_split_k_and_wide_layers_do_not():
    assert splits_for(64 * 4 * 4, 512, 9 * 512, 128) == 4            # layer 4 of ResNet-18/CIFAR at batch 64: 32 tiles, 36 slices
    assert splits_for(64 * 32 * 32, 64, 9 * 64, 128) == 1             # layer 1: 512 tiles


def _tma_im2col_model(x_nhwc, box_pixels, m0, tap, c0, cpb, KW, stride, pad, dil, OH, OW):
    """What one im2col-mode TMA load of the producer warp returns, modelled from the descriptor semantics the kernel
    relies on: start at base pixel (q*stride - pad, p*stride - pad, n) of output position m0, walk q -> p -> n, add the
    tap offset, zero outside the tensor."""
    B, H, W, C = x_nhwc.shape
    out = np.zeros((box_pixels, cpb), dtype=x_nhwc.dtype)
    r, s = tap // KW, tap % KW
    for i in range(box_pixels):
        m = m0 + i
        n, rem = divmod(m, OH * OW)
        p, q = divmod(rem, OW)
        h, w = p * stride - pad + r * dil, q * stride - pad + s * dil
        if n < B and 0 <= h < H and 0 <= w < W:
            out[i] = x_nhwc[n, h, w, c0:c0 + cpb]
    return out


def test_producer_coordinates_enumerate_the_im2col_matrix():
    rng = np.random.default_rng(0)
    for (B, C, H, W, k, stride, pad, dil) in ((2, 8, 6, 7, 3, 1, 1, 1), (3, 8, 9, 9, 3, 2, 1, 1), (2, 16, 8, 8, 1, 2, 0, 1), (1, 8, 9, 8, 3, 1, 2, 2)):
        x = rng.standard_normal((B, C, H, W)).astype(np.float32)
        OH = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
        OW = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
        cols = F.unfold(torch.from_numpy(x), (k, k), dilation=dil, padding=pad, stride=stride)             # [B, C*k*k, L]
        cols = cols.view(B, C, k * k, OH * OW).permute(0, 3, 2, 1).reshape(B * OH * OW, k * k * C).numpy()  # (kh, kw, c) order
        x_nhwc = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
        cpb, box = 8, 16                                                                                   # scaled-down channel block / tile
        M = B * OH * OW
        for m0 in range(0, M, box):
            for kb in range(k * k * (C // cpb)):
                tap, cb = divmod(kb, C // cpb)
                got = _tma_im2col_model(x_nhwc, box, m0, tap, cb * cpb, cpb, k, stride, pad, dil, OH, OW)
                rows = min(box, M - m0)
                assert np.array_equal(got[:rows], cols[m0:m0 + rows, kb * cpb:(kb + 1) * cpb])
                assert not got[rows:].any()
