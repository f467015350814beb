"""float_quantize (C++ CPU op) against the independent numpy oracle — SURVEY §4 row 1,
and BASELINE config 1 (CPU float_quantize(4,3) correctness, world_size=1)."""
import math

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from cpd_b200 import float_quantize
from oracle import quantize_ref_np, quantize_ieee_py

FORMATS = [(4, 3), (5, 2), (3, 0), (5, 10), (8, 7), (8, 23), (2, 1), (8, 0), (6, 9), (7, 16), (4, 23), (1, 2)]


def _bits_to_f32(bits: np.ndarray) -> np.ndarray:
    return bits.astype(np.uint32).view(np.float32)


def _assert_bit_equal(got: torch.Tensor, want: np.ndarray):
    g = got.numpy().view(np.uint32)
    w = want.view(np.uint32)
    nan_g, nan_w = np.isnan(got.numpy()), np.isnan(want)
    assert (nan_g == nan_w).all()
    bad = (g != w) & ~nan_w
    assert not bad.any(), f"{bad.sum()} mismatches, first at {np.flatnonzero(bad)[:5]}"


@pytest.mark.parametrize("e,m", FORMATS)
def test_exhaustive_bf16_grid(e, m):
    """All 2^16 fp32 patterns whose low 16 bits are zero (every sign/exponent, 7 mantissa bits)."""
    x = _bits_to_f32(np.arange(1 << 16, dtype=np.uint32) << 16)
    got = float_quantize(torch.from_numpy(x.copy()), e, m)
    _assert_bit_equal(got, quantize_ref_np(x, e, m))


@pytest.mark.parametrize("e,m", FORMATS)
def test_random_bit_patterns(e, m):
    rng = np.random.default_rng(1234 + 31 * e + m)
    x = _bits_to_f32(rng.integers(0, 1 << 32, size=1 << 18, dtype=np.uint64).astype(np.uint32))
    got = float_quantize(torch.from_numpy(x.copy()), e, m)
    _assert_bit_equal(got, quantize_ref_np(x, e, m))


@settings(max_examples=300, deadline=None)
@given(bits=st.integers(min_value=0, max_value=(1 << 32) - 1), e=st.integers(1, 8), m=st.integers(0, 23))
def test_hypothesis_any_format(bits, e, m):
    x = _bits_to_f32(np.array([bits], dtype=np.uint32))
    got = float_quantize(torch.from_numpy(x.copy()), e, m)
    _assert_bit_equal(got, quantize_ref_np(x, e, m))


@pytest.mark.parametrize("e,m", [(4, 3), (5, 2), (3, 0), (5, 10), (8, 7)])
def test_idempotent_monotone_symmetric(e, m):
    rng = np.random.default_rng(7)
    x = np.sort((rng.standard_normal(1 << 16) * np.exp2(rng.integers(-12, 12, 1 << 16))).astype(np.float32))
    t = torch.from_numpy(x)
    q = float_quantize(t, e, m)
    top = 2.0 ** ((1 << e) - 1 - ((1 << (e - 1)) - 1))          # quirk-A value 2^(emax+1): Q(top) = inf by design
    ok = q.abs() != top
    assert torch.equal(float_quantize(q, e, m)[ok], q[ok])                                                  # idempotent
    assert (q[1:] >= q[:-1]).all()                                                                          # monotone
    assert torch.equal(float_quantize(-t, e, m), -q)                                                       # odd


def test_known_values_e4m3():
    """(4,3): max normal 240 (all-ones exponent reserved), min subnormal 2^-9, quirk A at 248..256."""
    x = torch.tensor([240.0, 247.9, 248.0, 255.9, 256.0, 2.0 ** -9, 2.0 ** -10, 3 * 2.0 ** -11, 1.0625, 1.1875, 0.0, -0.0])
    q = float_quantize(x, 4, 3)
    want = [240.0, 240.0, 256.0, 256.0, math.inf, 2.0 ** -9, 0.0, 2.0 ** -9, 1.0, 1.25, 0.0, -0.0]
    assert q.tolist() == want
    assert math.copysign(1, q[-1].item()) == -1.0


def test_quirk_b_truncating_subnormal_shift():
    """0.5*ulp_min*(1+2^-20) -> 0 in reference mode (true RNE gives ulp_min)."""
    ulp = 2.0 ** -9
    v = torch.tensor([0.5 * ulp * (1 + 2.0 ** -20)], dtype=torch.float32)
    assert float_quantize(v, 4, 3).item() == 0.0
    assert float_quantize(v, 4, 3, rounding="ieee").item() == ulp


def test_specials_and_denormals():
    x = torch.tensor([float("inf"), float("-inf"), float("nan"), 1e-40, -1e-40])
    q = float_quantize(x, 5, 2)
    assert q[0] == math.inf and q[1] == -math.inf and math.isnan(q[2].item())
    assert q[3].item() == 0.0 and q[4].item() == 0.0 and math.copysign(1, q[4].item()) == 1.0   # sign dropped
    # (8,23) is the identity except for fp32 denormals
    y = torch.randn(1000)
    assert torch.equal(float_quantize(y, 8, 23), y)


@pytest.mark.parametrize("e,m", [(4, 3), (5, 2), (3, 0), (5, 10), (8, 7), (6, 9)])
def test_ieee_mode_matches_scalar_model(e, m):
    rng = np.random.default_rng(99)
    x = (rng.standard_normal(4000) * np.exp2(rng.integers(-40, 20, 4000))).astype(np.float32)
    got = float_quantize(torch.from_numpy(x), e, m, rounding="ieee").numpy()
    for xi, gi in zip(x.tolist(), got.tolist()):
        want = np.float32(quantize_ieee_py(xi, e, m))
        assert gi == want or (math.isinf(gi) and math.isinf(want) and gi * want > 0), (xi, gi, want)


def test_out_of_place_default_and_inplace_option():
    x = torch.randn(100)
    x0 = x.clone()
    q = float_quantize(x, 4, 3)
    assert torch.equal(x, x0) and q.data_ptr() != x.data_ptr()
    r = float_quantize(x, 4, 3, inplace=True)
    assert r.data_ptr() == x.data_ptr() and torch.equal(r, q)


def test_noncontiguous_zero_size_and_errors():
    x = torch.randn(33, 17).t()
    q = float_quantize(x, 5, 2)
    assert torch.equal(q, float_quantize(x.contiguous(), 5, 2))
    assert float_quantize(torch.empty(0), 4, 3).numel() == 0
    with pytest.raises(TypeError):
        float_quantize(torch.zeros(4, dtype=torch.float64), 4, 3)
    with pytest.raises(ValueError):
        float_quantize(torch.zeros(4), 9, 3)
    with pytest.raises(ValueError):
        float_quantize(torch.zeros(4), 4, 24)


def test_fused_shift_equals_scale_then_quantize():
    x = torch.randn(5000) * 1e-3
    for s in (-7, 0, 5, 11):
        assert torch.equal(float_quantize(x, 4, 3, shift=s), float_quantize(x * 2.0 ** s, 4, 3))


def test_wire_roundtrip_is_lossless_on_format_members():
    from cpd_b200.ops import native
    for (e, m) in [(4, 3), (5, 2), (3, 0), (5, 10), (8, 7), (2, 1), (6, 1)]:
        rng = np.random.default_rng(e * 100 + m)
        x = _bits_to_f32(rng.integers(0, 1 << 32, size=1 << 16, dtype=np.uint64).astype(np.uint32))
        q = float_quantize(torch.from_numpy(x), e, m)
        r = native.wire_roundtrip_cpu(q, e, m)
        assert torch.equal(r.nan_to_num(7.0), q.nan_to_num(7.0)), (e, m)
        assert native.wire_bits(e, m) in (8, 16, 32)
    assert native.wire_bits(4, 3) == 8 and native.wire_bits(5, 2) == 8 and native.wire_bits(5, 10) == 16
    assert native.wire_bits(8, 7) == 16 and native.wire_bits(8, 23) == 32 and native.wire_bits(7, 0) == 16 and native.wire_bits(3, 0) == 8
