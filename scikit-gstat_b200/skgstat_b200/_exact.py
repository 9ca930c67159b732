"""Host-side exact helpers: squared-space thresholds and bucket arithmetic.

np.digitize(d, edges) with d = fl(sqrt(s)) is reproduced on the device by
comparing s against T(e) = min{s : fl(sqrt(s)) >= e} (IEEE sqrt is monotone and
correctly rounded), see include/skg_b200.h.  All functions here are O(n_lags).
"""
from __future__ import annotations

import math
import struct

import numpy as np

INF_BITS = 0x7FF0000000000000


def to_bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def from_bits(b: int) -> float:
    return struct.unpack("<d", struct.pack("<Q", int(b)))[0]


def sq_threshold(e: float) -> float:
    """T(e) = smallest s >= 0 with fl(sqrt(s)) >= e."""
    e = float(e)
    if math.isnan(e):
        raise ValueError("bin edges contain NaN")
    if e <= 0.0:
        return 0.0
    if math.isinf(e):
        return math.inf
    s = e * e
    if math.isinf(s):
        return math.inf
    while s > 0.0 and math.sqrt(s) >= e:
        s = math.nextafter(s, -math.inf)
    while math.sqrt(s) < e:
        s = math.nextafter(s, math.inf)
    return s


def sq_threshold_bits(edges) -> np.ndarray:
    """uint64 bit patterns of T(e_k) for the (upper) lag-class edges."""
    return np.array([to_bits(sq_threshold(e)) for e in np.asarray(edges, dtype=float)],
                    dtype=np.uint64)


def sq_upper_inclusive_bits(limit: float) -> int:
    """Exclusive bit bound hi such that bits(s) < hi  <=>  fl(sqrt(s)) <= limit."""
    if math.isinf(limit):
        return INF_BITS + 1  # everything finite and +inf
    return to_bits(sq_threshold(math.nextafter(float(limit), math.inf)))


def bucket_of(v: float, lo: float, scale: float, n_buckets: int) -> int:
    """Device bucket function: min(B-1, trunc(fl(fl(v - lo) * scale))) (no FMA)."""
    t = (v - lo) * scale  # two IEEE roundings, as __dsub_rn/__dmul_rn
    if t >= 9.0e18:
        return n_buckets - 1
    return min(n_buckets - 1, int(t))


def bucket_lower_bits(b: int, lo_bits: int, hi_bits: int, scale: float, n_buckets: int) -> int:
    """Smallest bit pattern v in [lo_bits, hi_bits] with bucket_of(v) >= b (hi_bits if none).

    bucket_of is monotone in v, so this is a bisection; it starts from the analytic boundary
    lo + b / scale, brackets the true one (a few ulps away: two roundings) by doubling steps and
    only bisects inside that bracket - a handful of evaluations instead of ~60."""
    lo_val = from_bits(lo_bits)

    def reached(bits):
        return bucket_of(from_bits(bits), lo_val, scale, n_buckets) >= b

    a, z = lo_bits, hi_bits
    if z - a > 64 and scale > 0.0 and math.isfinite(scale):
        guess = lo_val + b / scale
        if math.isfinite(guess) and guess >= 0.0:
            g = min(max(to_bits(guess), a), z - 1)
            step = 1
            if reached(g):
                z = g
                while z - step >= a and reached(z - step):
                    z -= step
                    step *= 2
                a = max(a, z - step + 1)
            else:
                a = g + 1
                while a - 1 + step < z and not reached(a - 1 + step):
                    a += step
                    step *= 2
                z = min(z, a - 1 + step)
    while a < z:
        mid = (a + z) // 2
        if reached(mid):
            z = mid
        else:
            a = mid + 1
    return a
