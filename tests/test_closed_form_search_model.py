"""A numpy / exact-rational model of the closed-form systematic-resampling search used by the kernels
(count_positions_below and its float32 pre-filter in aesmc_b200/csrc/common.cuh), checked on the CPU against the
reference's own expression: pos = (u + arange(K)) / K in float64, np.digitize against the float32 CDF
(inference.py:251,264).  For particle j the kernels need c_j = #{k : pos_k < cdfn_j}:

    float64 form   t = fma(c, K, -u); if |t - rint(t)| > K * 2^-50: ceil(t), else the reference's expression near rint(t)
    float32 filter tf = fma32(c, K, -u32); if |tf - rint(tf)| > K * 2^-23 + 2^-24: ceil(tf), else the float64 form
"""
import math
from fractions import Fraction

import numpy as np
import pytest

f32 = np.float32


def rn(fr, mant_bits):
    """Round an exact rational to the nearest binary float with mant_bits of precision (ties to even); normal range."""
    if fr == 0:
        return Fraction(0)
    sign = -1 if fr < 0 else 1
    a = abs(fr)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if Fraction(2) ** e > a:
        e -= 1
    unit = Fraction(2) ** (e - (mant_bits - 1))
    q = a / unit
    n = q.numerator // q.denominator
    rem = q - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and n & 1):
        n += 1
    return sign * n * unit


def reference_count(c, u, K):
    pos = (u + np.arange(K)) / K          # float64, as inference.py:251
    return int(np.searchsorted(pos, np.float64(c), side="left"))   # #{k : pos_k < c}


def count_f64(c, u, K):
    t = float(rn(Fraction(float(c)) * K - Fraction(u), 53))        # fma(c, K, -u) in float64
    r = float(np.rint(t))
    if abs(t - r) > K * 8.8817841970012523e-16:
        return min(max(int(math.ceil(t)), 0), K)
    k = min(max(int(r) - 1, 0), K)
    while k < K and (u + k) / K < float(c):
        k += 1
    while k > 0 and not ((u + (k - 1)) / K < float(c)):
        k -= 1
    return k


def count_f32_filtered(c, u, K):
    u32 = f32(u)
    tf = f32(float(rn(Fraction(float(c)) * K - Fraction(float(u32)), 24)))   # fma32(c, K, -u32)
    tol32 = f32(K) * f32(2.0 ** -23) + f32(2.0 ** -24)
    r = np.rint(tf)
    if abs(tf - r) > tol32:
        return min(int(r) + (1 if tf - r > 0 else 0), K), True
    return count_f64(c, u, K), False


@pytest.mark.parametrize("K", [1, 2, 3, 100, 1000, 4096, 16384, 65536])
def test_closed_form_equals_digitize(K):
    rng = np.random.default_rng(K)
    fast = total = 0
    for trial in range(400):
        u = float(rng.random())
        if trial % 4 == 0:      # CDF entries sitting on (or one float32 ulp around) a stratified position
            k = int(rng.integers(0, K))
            c = f32((u + k) / K)
            c = [c, np.nextafter(c, f32(0)), np.nextafter(c, f32(2))][trial % 3]
        elif trial % 4 == 1:    # ... and the float32 neighbours of k / K
            c = np.nextafter(f32(rng.integers(0, K + 1) / K), f32(rng.integers(0, 2) * 2))
        else:
            c = f32(rng.random())
        c = f32(min(max(float(c), 0.0), 1.0))
        want = reference_count(c, u, K)
        assert count_f64(c, u, K) == want, (K, u, float(c))
        if K < (1 << 20):
            got, decided = count_f32_filtered(c, u, K)
            assert got == want, (K, u, float(c), decided)
            if trial % 4 >= 2:   # the random entries only: the adversarial ones are built to need float64
                fast += decided
                total += 1
    if K <= 4096:
        assert fast > 0.9 * total   # the float32 filter decides almost everything at these sizes
