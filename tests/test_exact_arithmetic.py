"""A third opinion on the arithmetic, from first principles.

DESIGN.md section 3 DEFINES the score (the reference has no arithmetic of its own): fp32 fused
multiply-adds in a fixed lane-strided order, a fixed addition tree, correctly rounded sqrt and
divide, zero for degenerate denominators.  oracle/oracle.cpp implements that definition with
compiler intrinsics and flags (-ffp-contract=off, __builtin_fmaf); the kernels implement it with
__fmaf_rn & co.  This file implements it a third time with NO floating-point arithmetic at all:
every operation is computed exactly in rational numbers (fractions.Fraction) and rounded to
binary32 (or bfloat16) by an explicit round-to-nearest-even on integers.  If the oracle's bits
equal these, the oracle computes what the specification says -- not merely what its compiler
happened to emit -- and the GPU parity tests (bit-exact against the oracle) inherit that.
"""
from fractions import Fraction
from math import isqrt

import numpy as np
import pytest

F32_MAX = Fraction((2 ** 24 - 1) * 2 ** 104)  # largest finite binary32


def _exact(x) -> Fraction:
    return Fraction(float(np.float32(x)))  # binary32 -> binary64 -> rational: both exact


def _round(fr: Fraction, mant_bits: int = 24) -> Fraction:
    """round-to-nearest-even of a rational to a binary float with `mant_bits` significant bits
    and binary32's exponent range (subnormals included); +-inf are returned as None."""
    if fr == 0:
        return Fraction(0)
    sign = -1 if fr < 0 else 1
    a = abs(fr)
    # e = floor(log2(a))
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if Fraction(2) ** e > a:
        e -= 1
    assert Fraction(2) ** e <= a < Fraction(2) ** (e + 1)
    q = max(e - (mant_bits - 1), -149)  # exponent of the rounding quantum
    scaled = a / Fraction(2) ** q
    n = scaled.numerator // scaled.denominator
    rem = scaled - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and n % 2 == 1):
        n += 1
    val = n * Fraction(2) ** q
    if val > F32_MAX:
        return None
    return sign * val


# None stands for "not finite" (an overflowed sum, and anything computed from one): every way
# such a value can reach the score -- dot/inf = 0, inf/x = inf, inf/inf = NaN -- ends in the
# rule's 0.0, so it only has to propagate.
def _fma(a, b, c):
    return None if a is None or b is None or c is None else _round(a * b + c)


def _add(a, b):
    return None if a is None or b is None else _round(a + b)


def _mul(a, b):
    return None if a is None or b is None else _round(a * b)


def _div(a, b):
    return None if a is None or b is None else _round(a / b)


def _sqrt(x: Fraction) -> Fraction:
    """correctly rounded binary32 square root of a non-negative binary32 value"""
    if x == 0:
        return Fraction(0)
    # scale to an integer with plenty of bits: x = X / 4^s  ->  sqrt(x) = sqrt(X) / 2^s
    s = 200
    X = x * Fraction(4) ** s
    assert X.denominator == 1
    r = isqrt(X.numerator)
    lo = Fraction(r, 2 ** s)  # lo <= sqrt(x) < lo + 2^-s: far finer than a binary32 ulp
    cand = _round(lo)
    # sqrt of a binary32 is never a rounding midpoint, so the correctly rounded result is the
    # binary32 value whose rounding interval (between the midpoints to its neighbours) holds
    # sqrt(x); rounding `lo` can only miss it when lo and sqrt(x) straddle a midpoint, so it is
    # cand or one of its neighbours -- decided exactly, by comparing squares
    for z in (cand, _next_up(cand), _next_dn(cand)):
        if z <= 0:
            continue
        half_up = z + (_next_up(z) - z) / 2
        half_dn = z - (z - _next_dn(z)) / 2
        if half_dn * half_dn < x < half_up * half_up:
            return z
    raise AssertionError("sqrt rounding search failed")


def _next_up(z: Fraction) -> Fraction:
    e = z.numerator.bit_length() - z.denominator.bit_length()
    if Fraction(2) ** e > z:
        e -= 1
    return z + Fraction(2) ** max(e - 23, -149)


def _next_dn(z: Fraction) -> Fraction:
    e = z.numerator.bit_length() - z.denominator.bit_length()
    if Fraction(2) ** e > z:
        e -= 1
    step = Fraction(2) ** max(e - 23, -149)
    if z - step < Fraction(2) ** e and e - 24 >= -149:  # crossing a binade: the ulp below halves
        step = Fraction(2) ** (e - 24)
    return z - step


def _canon_dot(x, y):
    """sum x_j y_j in the canonical order of DESIGN.md section 3 (x, y: lists of Fractions whose
    length is a multiple of 128)"""
    stripes = len(x) // 128
    part = []
    for lane in range(32):
        acc = [Fraction(0)] * 4
        for s in range(stripes):
            for c in range(4):
                j = 128 * s + 4 * lane + c
                acc[c] = _fma(x[j], y[j], acc[c])
        part.append(_add(_add(acc[0], acc[1]), _add(acc[2], acc[3])))
    for m in (16, 8, 4, 2, 1):
        part = [_add(part[l], part[l ^ m]) for l in range(32)]
    return part[0]


def _to_bits(fr: Fraction) -> int:
    return int(np.float32(float(fr)).view(np.uint32))  # fr is a binary32 value: exact


def _score_bits(row, q, bf16=False):
    dim = len(row)
    padded = ((dim + 127) // 128) * 128
    e = [_exact(v) for v in row]
    if bf16:  # storage rounding: round-to-nearest-even to 8 significant bits
        e = [_round(v, 8) for v in e]
    e += [Fraction(0)] * (padded - dim)
    qq = [_exact(v) for v in q] + [Fraction(0)] * (padded - dim)
    dot, ne2, nq2 = _canon_dot(qq, e), _canon_dot(e, e), _canon_dot(qq, qq)
    if dot is None or ne2 is None or nq2 is None:
        return 0  # an overflowing sum: the rule scores it 0.0
    den = _mul(_sqrt(nq2), _sqrt(ne2))  # (nq2, ne2 are finite here)
    if den is None or not den > 0:
        return 0
    s = _div(dot, den)
    if s is None:
        return 0
    return _to_bits(s)  # (-0.0 cannot arise: a zero dot with den > 0 rounds to +0 here)


@pytest.mark.parametrize("dim", [1, 3, 100, 128, 384, 1000])
def test_oracle_scores_equal_exact_rational_arithmetic(orc, dim):
    rng = np.random.default_rng(dim)
    rows = rng.standard_normal((7, dim)).astype(np.float32)
    rows[1] *= 1e-3
    rows[2] *= 300.0
    rows[3] = np.abs(rows[3])             # no cancellation
    rows[4] = 0.0                          # zero row -> 0.0
    rows[5, ::2] = 0.0
    rows[6] = rows[0] * np.float32(1.0000001)
    qs = rng.standard_normal((2, dim)).astype(np.float32)
    qs[1] = rows[0] * np.float32(3.0)      # parallel to a row: score rounds near 1
    for q in qs:
        got = orc.scores(rows, q).view(np.uint32)
        want = [_score_bits(r, q) for r in rows]
        assert [int(g) for g in got] == want


@pytest.mark.parametrize("dim", [3, 384])
def test_oracle_bf16_scores_equal_exact_rational_arithmetic(orc, dim):
    rng = np.random.default_rng(100 + dim)
    rows = (rng.standard_normal((5, dim)) * rng.uniform(0.01, 50.0)).astype(np.float32)
    q = rng.standard_normal(dim).astype(np.float32)
    got = orc.scores(rows, q, bf16=True).view(np.uint32)
    want = [_score_bits(r, q, bf16=True) for r in rows]
    assert [int(g) for g in got] == want


def test_degenerate_inputs_follow_the_rule(orc):
    dim = 384
    rows = np.zeros((3, dim), np.float32)
    rows[1, 0] = 1e-30                     # norm underflows towards the subnormals
    rows[2] = 1e19                         # sum of squares overflows binary32
    q = np.ones(dim, np.float32)
    got = orc.scores(rows, q).view(np.uint32)
    assert [int(g) for g in got] == [_score_bits(r, q) for r in rows]
    zero_q = np.zeros(dim, np.float32)
    assert [int(g) for g in orc.scores(rows, zero_q).view(np.uint32)] == [0, 0, 0]


def test_rounding_helper_against_numpy():
    """the rational rounding itself, against numpy's binary64 -> binary32 conversion (which is a
    single correctly rounded step for binary64 inputs)"""
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.standard_normal(200) * 10.0 ** rng.integers(-40, 38, 200),
                         [1.0 + 2.0 ** -24, 1.0 + 3 * 2.0 ** -24, 2.0 ** -149 * 0.5, 2.0 ** -149 * 1.5,
                          2.0 ** -126 * (1 - 2.0 ** -25), 3.4028235677973366e38, 3.5e38]])
    for x in xs:
        r = _round(Fraction(float(x)))
        with np.errstate(over="ignore"):
            want = np.float32(x)
        if np.isinf(want):
            assert r is None
        else:
            assert r == Fraction(float(want)), x
    for x in np.abs(rng.standard_normal(100).astype(np.float32)) * np.float32(10.0) ** rng.integers(-18, 18, 100):
        x = np.float32(x)
        assert _sqrt(_exact(x)) == _exact(np.sqrt(x)), x   # IEEE sqrt is correctly rounded
