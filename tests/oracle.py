"""Pure-python/numpy oracles of the CPD numerics, written from the spec —
independent of the C++/CUDA implementations they are used to check."""
from __future__ import annotations

import math

import numpy as np


def quantize_ref_np(x: np.ndarray, exp_bits: int, man_bits: int) -> np.ndarray:
    """Vectorised bit-level model of the reference cast (quirks A, B, C included).

    Follows the reference's own control flow: classify by the fp32 exponent, overflow test
    on the un-rounded exponent, three-way RNE on the 24-bit significand, truncating
    shift for target subnormals, value rebuilt as man/2^23 * 2^e."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.int64)
    sign = (u >> 31) & 1
    E = (u >> 23) & 0xFF
    f = u & 0x7FFFFF
    out = np.empty(x.shape, dtype=np.float64)

    bias = (1 << (exp_bits - 1)) - 1
    passthru = (E == 255) | ((E == 0) & (f == 0))
    denorm = (E == 0) & (f != 0)
    normal = ~passthru & ~denorm

    man = f | (1 << 23)
    true_exp = E - 127
    new_e = true_exp + bias
    overflow = normal & (new_e >= (1 << exp_bits) - 1)
    tnorm = normal & ~overflow & (new_e > 0)
    tsub = normal & ~overflow & (new_e <= 0)

    def rne(m, skip_m23):
        drop = 23 - man_bits
        if drop == 0:
            return m.copy()                      # (m==23): normal branch short-circuits; subnormal: 1<<-1 == 0 -> keep
        rb = (m >> (drop - 1)) & 1
        sticky = m & ((1 << (drop - 1)) - 1)
        lsb = (m >> drop) & 1
        up = (rb == 1) & ((sticky != 0) | (lsb == 1))
        m2 = np.where(up, m + (1 << (drop - 1)), m)
        return m2 & ~((1 << drop) - 1)

    man_n = rne(man, True)
    val_n = man_n.astype(np.float64) / float(1 << 23) * np.exp2(true_exp.astype(np.float64))

    shift = np.clip(1 - new_e, 0, 40)
    man_s = np.where(shift >= 32, 0, man >> np.minimum(shift, 31))
    man_s = rne(man_s, False)
    val_s = man_s.astype(np.float64) / float(1 << 23) * (2.0 ** (1 - bias))

    out[:] = 0.0
    res = np.where(tnorm, val_n, np.where(tsub, val_s, 0.0))
    res = np.where(sign == 1, -res, res)
    with np.errstate(over="ignore"):
        res32 = res.astype(np.float32)
    res32 = np.where(overflow, np.where(x > 0, np.float32(np.inf), np.float32(-np.inf)), res32)
    res32 = np.where(denorm, np.float32(0.0), res32)
    res32 = np.where(passthru, x, res32)
    return res32.astype(np.float32)


def quantize_ieee_py(v: float, exp_bits: int, man_bits: int) -> float:
    """Scalar IEEE-754 RNE to (exp_bits, man_bits) with subnormals; overflow after rounding."""
    if math.isnan(v) or math.isinf(v) or v == 0.0:
        return v
    bias = (1 << (exp_bits - 1)) - 1
    emin = 1 - bias
    emax = (1 << exp_bits) - 2 - bias
    a = abs(v)
    e = math.floor(math.log2(a))
    if 2.0 ** e > a:
        e -= 1
    if 2.0 ** (e + 1) <= a:
        e += 1
    e = max(e, emin)
    q = 2.0 ** (e - man_bits)
    n = a / q
    r = math.floor(n)
    diff = n - r
    if diff > 0.5 or (diff == 0.5 and r % 2 == 1):
        r += 1
    res = r * q
    if res >= 2.0 ** (emax + 1):
        res = math.inf
    return math.copysign(res, v)


def q_scalar(v, e, m):
    return float(quantize_ref_np(np.array([v], dtype=np.float32), e, m)[0])


def gemm_exact_oracle(A: np.ndarray, B: np.ndarray, e: int, m: int, kahan: bool = True) -> np.ndarray:
    """Sequential-k, per-MAC quantised (Kahan) accumulation — SURVEY §2.9.2, vectorised over (i, j)."""
    A = A.astype(np.float32)
    B = B.astype(np.float32)
    M, K = A.shape
    N = B.shape[1]
    Q = lambda t: quantize_ref_np(t.astype(np.float32), e, m)  # noqa: E731
    s = np.zeros((M, N), np.float32)
    c = np.zeros((M, N), np.float32)
    for k in range(K):
        p = Q(np.outer(A[:, k], B[k, :]).astype(np.float32))
        if kahan:
            y = Q(p - c)
            t = Q(s + y)
            c = Q(Q(t - s) - y)
            s = t
        else:
            s = Q(s + p)
    return s


def aps_fold_oracle(grads, e: int, m: int, use_aps: bool, kahan: bool, world_mult: float):
    """grads: list (ranks) of 1-D float32 arrays of ONE tensor.  Returns (result, shift)."""
    Q = lambda t: quantize_ref_np(np.asarray(t, dtype=np.float32), e, m)  # noqa: E731
    shift = 0
    if use_aps:
        me = -math.inf
        for g in grads:
            amax = np.float32(np.max(np.abs(g))) if g.size else np.float32(0)
            v = np.float32(amax * np.float32(world_mult))
            mr = math.ceil(math.log2(float(v))) if v > 0 else -math.inf
            me = max(me, mr)
        bias = (1 << (e - 1)) - 1
        shift = 0 if not math.isfinite(me) else int(max(-126, min(126, bias - me)))
    sc = np.float32(2.0 ** shift)
    res = np.zeros_like(grads[0], dtype=np.float32)
    c = np.zeros_like(res)
    ident = (e == 8 and m == 23 and not kahan)
    for g in grads:
        g = g.astype(np.float32)
        if use_aps:
            g = Q(g * sc)
        if ident:
            res = (res + g).astype(np.float32)
        elif kahan:
            y = Q(g - c)
            t = Q(res + y)
            c = Q(Q(t - res) - y)
            res = t
        else:
            res = Q(res + g)
    if use_aps:
        res = (res * np.float32(2.0 ** -shift)).astype(np.float32)
    return res, shift
