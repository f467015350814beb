"""Bit-for-bit parity with the UNMODIFIED reference extension (built from /root/reference into
baseline/_ref) — SURVEY §4 'Parity vs the reference (GPU)'."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref_quant():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "CPDtorch")):
        pytest.skip("baseline/_ref not installed")
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    from prebuild_ref_ext import ref_env
    ref_env()
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_quant_function", os.path.join(ref, "CPDtorch", "quant", "quant_function.py"))
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"reference extension does not build here: {e}")
    return mod


@pytest.mark.parametrize("e,m", [(4, 3), (5, 2), (3, 0), (5, 10), (8, 7), (8, 23), (6, 9)])
def test_float_quantize_bitwise_vs_reference(ref_quant, e, m):
    from cpd_b200 import float_quantize
    rng = np.random.default_rng(e * 100 + m)
    x = torch.from_numpy(rng.integers(0, 1 << 32, size=1 << 20, dtype=np.uint64).astype(np.uint32).view(np.float32).copy()).cuda()
    ours = float_quantize(x, e, m)
    theirs = ref_quant.float_quantize(x.clone(), e, m)         # the reference mutates its input
    same = (ours.view(torch.int32) == theirs.view(torch.int32)) | (ours.isnan() & theirs.isnan())
    assert bool(same.all()), f"{int((~same).sum())} mismatches"


@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (128, 48, 100)])
def test_quant_gemm_bitwise_vs_reference(ref_quant, M, N, K):
    from cpd_b200 import quant_gemm
    torch.manual_seed(0)
    a, b = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
    for (e, m) in [(4, 3), (5, 2), (8, 23)]:
        ours = quant_gemm(a, b, man=m, exp=e, mode="exact")
        theirs = ref_quant.quant_gemm(a, b, man=m, exp=e)      # M % 16 == 0: the reference's sound branch
        assert torch.equal(ours, theirs), (e, m)
