"""quant_gemm / Quant_Linear / Quant_Conv on CPU vs the python oracle (SURVEY §4 row 2;
BASELINE config 1: 64x64 GEMM correctness on CPU, world_size=1)."""
import numpy as np
import pytest
import torch

from cpd_b200 import quant_gemm, Quant_Linear, Quant_Conv, Quantizer, float_quantize
from cpd_b200.ops import native
from oracle import gemm_exact_oracle


def _eq(a: torch.Tensor, b: np.ndarray):
    an = a.numpy()
    same = (an.view(np.uint32) == b.view(np.uint32)) | (np.isnan(an) & np.isnan(b))
    assert same.all(), f"{(~same).sum()} mismatches"


@pytest.mark.parametrize("e,m", [(4, 3), (5, 2), (5, 10), (8, 23)])
def test_gemm_64x64_exact_matches_oracle(e, m):
    rng = np.random.default_rng(e * 10 + m)
    A = rng.standard_normal((64, 64)).astype(np.float32)
    B = rng.standard_normal((64, 64)).astype(np.float32)
    C = quant_gemm(torch.from_numpy(A), torch.from_numpy(B), man=m, exp=e, mode="exact")
    _eq(C, gemm_exact_oracle(A, B, e, m, kahan=True))


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (16, 17, 63), (17, 16, 65), (63, 64, 7), (65, 1, 16), (10, 512, 33)])
def test_gemm_ragged_shapes(M, N, K):
    rng = np.random.default_rng(M * 1000 + N * 10 + K)
    A = (rng.standard_normal((M, K)) * 0.5).astype(np.float32)
    B = (rng.standard_normal((K, N)) * 0.5).astype(np.float32)
    C = quant_gemm(torch.from_numpy(A), torch.from_numpy(B), man=2, exp=5, mode="exact")
    _eq(C, gemm_exact_oracle(A, B, 5, 2))


def test_gemm_plain_variant_and_strided_operands():
    rng = np.random.default_rng(3)
    A = rng.standard_normal((20, 30)).astype(np.float32)
    B = rng.standard_normal((30, 12)).astype(np.float32)
    C = native.gemm_exact(torch.from_numpy(A), torch.from_numpy(B), 4, 3, kahan=False)
    _eq(C, gemm_exact_oracle(A, B, 4, 3, kahan=False))
    # transposed views are consumed in place
    At = torch.from_numpy(np.ascontiguousarray(A.T)).t()
    Bt = torch.from_numpy(np.ascontiguousarray(B.T)).t()
    assert not At.is_contiguous() and not Bt.is_contiguous()
    _eq(native.gemm_exact(At, Bt, 4, 3), gemm_exact_oracle(A, B, 4, 3))


def test_gemm_fp32_format_close_to_mm():
    a, b = torch.randn(96, 200), torch.randn(200, 48)
    ref = (a.double() @ b.double()).float()
    assert (quant_gemm(a, b) - ref).abs().max() < 1e-4       # (8,23) + Kahan: better than plain fp32


def test_tile_cpu_model_error_bound():
    """The tile semantics (per-K-slice requant) differ from exact by design; check against fp64."""
    a, b = torch.randn(64, 512), torch.randn(512, 64)
    ref = (a.double() @ b.double()).float()
    for (e, m, tol) in [(5, 10, 2e-2), (8, 7, 0.15), (8, 23, 1e-4)]:
        c = native.gemm_tile_cpu(a, b, e, m, k_slice=128)
        rel = (c - ref).abs().max() / ref.abs().max()
        assert rel < tol, (e, m, rel)
    # one slice covering all of K, (8,23): equals the fp64-accumulated product rounded once
    c = native.gemm_tile_cpu(a, b, 8, 23, k_slice=512)
    assert torch.equal(c, ref)


def test_quant_linear_forward_backward():
    torch.manual_seed(0)
    lin = Quant_Linear(24, 10, bias=True, exp=5, man=10, mode="exact")
    x = torch.randn(6, 24, requires_grad=True)
    y = lin(x)
    W, b = lin.weight.detach().numpy(), lin.bias.detach().numpy()
    want = gemm_exact_oracle(x.detach().numpy(), W.T.copy(), 5, 10) + b[None, :]
    _eq(y.detach(), want.astype(np.float32))
    g = torch.randn_like(y)
    y.backward(g)
    _eq(x.grad, gemm_exact_oracle(g.numpy(), W, 5, 10))
    _eq(lin.weight.grad, gemm_exact_oracle(g.numpy().T.copy(), x.detach().numpy(), 5, 10))   # M = 10: ragged, must be right
    assert torch.equal(lin.bias.grad, float_quantize(g.sum(0), 5, 10))


def test_quant_linear_fp32_matches_nn_linear():
    torch.manual_seed(1)
    q = Quant_Linear(32, 16, exp=8, man=23, mode="exact")
    ref = torch.nn.Linear(32, 16)
    ref.load_state_dict(q.state_dict())
    x = torch.randn(5, 3, 32)
    assert torch.allclose(q(x), ref(x), atol=1e-5)


def test_quant_conv_matches_conv2d_in_fp32_and_shapes():
    torch.manual_seed(2)
    for (k, s, p, d) in [(3, 1, 1, 1), (3, 2, 1, 1), (1, 1, 0, 1), (3, 1, 2, 2)]:
        qc = Quant_Conv(4, 6, k, stride=s, padding=p, dilation=d, bias=True, exp=8, man=23, mode="exact")
        x = torch.randn(2, 4, 9, 9, requires_grad=True)
        y = qc(x)
        ref = torch.nn.functional.conv2d(x, qc.weight, qc.bias, stride=s, padding=p, dilation=d)
        assert y.shape == ref.shape
        assert torch.allclose(y, ref, atol=1e-4)
        y.sum().backward()
        assert qc.weight.grad is not None and x.grad is not None
    with pytest.raises(NotImplementedError):
        Quant_Conv(4, 4, 3, groups=2)


def test_quant_conv_low_precision_equals_unfold_gemm():
    torch.manual_seed(3)
    qc = Quant_Conv(3, 5, 3, padding=1, bias=False, exp=4, man=3, mode="exact")
    x = torch.randn(2, 3, 6, 6)
    y = qc(x)
    cols = torch.nn.functional.unfold(x, 3, padding=1).transpose(1, 2).reshape(-1, 27)
    want = gemm_exact_oracle(cols.numpy(), qc.weight.detach().view(5, -1).t().contiguous().numpy(), 4, 3)
    _eq(y.detach().view(2, 5, 36).transpose(1, 2).reshape(-1, 5).contiguous(), want)


def test_quantizer_module_forward_and_backward_casts():
    qz = Quantizer(4, 3, 5, 2)
    x = torch.randn(50, requires_grad=True)
    y = qz(x)
    assert torch.equal(y.detach(), float_quantize(x.detach(), 4, 3))
    g = torch.randn(50)
    y.backward(g)
    assert torch.equal(x.grad, float_quantize(g, 5, 2))
    ident = Quantizer()
    z = torch.randn(5, requires_grad=True)
    assert torch.equal(ident(z), z)


def test_tile_model_split_k_is_an_ordered_fold_of_the_splits():
    """splits=1 is the unsplit model; splits=S partitions the K-slices into S contiguous ranges, accumulates each from
    zero and folds the S results in order with a cast after every add."""
    from cpd_b200 import float_quantize
    torch.manual_seed(0)
    a, b = torch.randn(5, 512), torch.randn(512, 7)
    one = native.gemm_tile_cpu(a, b, 5, 2, rounding="ieee", k_slice=64, splits=1)
    assert torch.equal(one, native.gemm_tile_cpu(a, b, 5, 2, rounding="ieee", k_slice=64))
    four = native.gemm_tile_cpu(a, b, 5, 2, rounding="ieee", k_slice=64, splits=4)
    parts = [native.gemm_tile_cpu(a[:, z * 128:(z + 1) * 128].contiguous(), b[z * 128:(z + 1) * 128].contiguous(), 5, 2,
                                  rounding="ieee", k_slice=64) for z in range(4)]
    acc = torch.zeros_like(parts[0])
    for p in parts:
        acc = float_quantize(acc + p, 5, 2, "ieee")
    assert torch.equal(four, acc)
