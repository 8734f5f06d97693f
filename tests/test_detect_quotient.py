"""The quotient of demod.java:451, avg = ((float)k*avg + x)/(float)(k+1), as k_detect forms it
(java-sdr_b200/csrc/demod_fir.cu): with t = RN(k*avg), a = RN(t + x) and 1/n = y_hi + y_lo as a float
pair, q = fma(a, y_hi, fma(t, y_lo, RN(x*y_lo))).  Checked here in exact rational arithmetic
against the correctly rounded float quotient a/n (what Java's `/` and the oracle's C `/`
compute), on random operands and on the operands whose quotients lie closest to a rounding
boundary of the float format.  No GPU needed: this is the arithmetic argument, the GPU parity
test of the kernel is tests/test_gpu_detect.py."""
from fractions import Fraction

import numpy as np


def rn_f32(v: Fraction) -> Fraction:
    """Round a rational in the normal float range to the nearest float, ties to even."""
    if v == 0:
        return v
    if v < 0:
        return -rn_f32(-v)
    e = v.numerator.bit_length() - v.denominator.bit_length()      # 2^(e-1) < v < 2^(e+1)
    if Fraction(2) ** e > v:
        e -= 1
    assert Fraction(2) ** e <= v < Fraction(2) ** (e + 1) and e > -126
    ulp = Fraction(2) ** (e - 23)
    q = v / ulp                                                    # in [2^23, 2^24)
    m = q.numerator // q.denominator
    rem = q - m
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (m & 1)):
        m += 1
    return m * ulp


def F(x) -> Fraction:
    return Fraction(float(x))


def check(t: np.float32, x: np.float32, n: int):
    """t = RN(k*avg) >= 0, x >= 0: the kernel's q against RN(a/n), a = RN(t + x)."""
    t, x = np.float32(t), np.float32(x)
    a = t + x                                                      # __fadd_rn
    r = 1.0 / float(n)                                             # __drcp_rn
    yh = np.float32(r)                                             # __double2float_rn
    yl = np.float32(r - float(yh))                                 # exact difference, rounded once
    xyl = x * yl                                                   # __fmul_rn, filled in per tile
    p = rn_f32(F(t) * F(yl) + F(xyl))                              # __fmaf_rn(t, y_lo, x*y_lo)
    got = rn_f32(F(a) * F(yh) + p)                                 # __fmaf_rn(a, y_hi, p)
    want = rn_f32(F(a) / n)
    assert got == want, (float(t), float(x), n, float(got), float(want))
    # and the host's own float division agrees with the rational rounding
    assert F(a / np.float32(n)) == want


def splits(a: np.float32, rng):
    """(t, x) with t, x >= 0 and RN(t + x) == a."""
    yield a, np.float32(0)
    yield np.float32(0), a
    yield a / np.float32(2), a / np.float32(2)
    for _ in range(3):
        t = np.float32(float(a) * rng.uniform(0, 1))
        x = a - t
        if x >= 0 and t + x == a:
            yield t, x


def test_random_operands():
    rng = np.random.default_rng(451)
    for _ in range(15000):
        n = int(rng.integers(1, 1 << 19))
        a = np.float32(rng.uniform(0.5, 1.0) * 2.0 ** int(rng.integers(-38, 85)))
        t = np.float32(float(a) * rng.uniform(0, 1))
        check(t, a - t, n)
    for n in (1, 2, 3, 5, 7, 1023, 1024, 1025, 19200, 131071, 131072, (1 << 19) - 1):
        for a in (1.0, 3.0, 1.5, 0.1, 16777215.0, 8388609.0, 2.0 ** -39, 2.0 ** 85):
            for t, x in splits(np.float32(a), rng):
                check(t, x, n)


def test_quotients_next_to_rounding_boundaries():
    """a/n closest to a midpoint between two floats: a = the floats around n * (odd 25-bit M)."""
    rng = np.random.default_rng(2)
    cases = 0
    for _ in range(4000):
        n = int(rng.integers(3, 1 << 19))
        M = int(rng.integers(1 << 24, 1 << 25)) | 1                # midpoint of [2^24, 2^25) in units of 1/2 ulp
        e = int(rng.integers(-70, 30))
        target = Fraction(n * M) * Fraction(2) ** e                # a/n would be exactly the midpoint
        a0 = np.float32(float(target))
        for a in (np.nextafter(a0, np.float32(0)), a0, np.nextafter(a0, np.float32(np.inf))):
            if 2.0 ** -39 <= float(a) <= 2.0 ** 85:
                for t, x in splits(np.float32(a), rng):
                    check(t, x, n)
                    cases += 1
    assert cases > 20000
