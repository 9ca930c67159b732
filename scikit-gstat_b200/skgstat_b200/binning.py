"""Lag-class edges.

``even_width_lags`` / ``uniform_count_lags`` keep the reference's operator
signature ``bin_func(distances, n, maxlag) -> (edges, None)`` (binning.py:10-79)
for callers that hold a materialised distance vector.  The Variogram classes
do NOT call them on the hot path: they build the same edges from a handful of
exact scalars delivered by the pair kernels (max distance, order statistics),
with numpy's own arithmetic, so the edges are bit-identical
(``even_edges`` / ``percentile_from_order_stats``).
"""
from __future__ import annotations

import numpy as np


def even_width_lags(distances, n, maxlag):
    """binning.py:10-40."""
    d = np.asarray(distances, dtype=float)
    if maxlag is None or maxlag > np.nanmax(d):
        maxlag = np.nanmax(d)
    return np.linspace(0, maxlag, n + 1)[1:], None


def uniform_count_lags(distances, n, maxlag):
    """binning.py:43-79."""
    d = np.asarray(distances, dtype=float)
    if maxlag is None or maxlag > np.nanmax(d):
        maxlag = np.nanmax(d)
    d = d[np.where(d <= maxlag)]
    return np.fromiter((np.nanpercentile(d, (i / n) * 100) for i in range(1, n + 1)),
                       dtype=float), None


# ---- scalar forms used with the device-delivered statistics -----------------

def even_edges(n, maxlag, dmax):
    """binning.py:37-40 with nanmax(distances) = dmax."""
    if maxlag is None or maxlag > dmax:
        maxlag = dmax
    return np.linspace(0, maxlag, n + 1)[1:]


def percentile_ranks(total: int, q: float):
    """0-based ranks np.percentile(a, q) (method 'linear') reads from sorted a of
    length `total`: floor((total-1)*q/100) and the next one
    (numpy/lib/_function_base_impl.py: _quantile, virtual index (n-1)*q)."""
    quant = np.true_divide(q, 100.0)
    vi = (total - 1) * quant
    prev = int(np.floor(vi))
    prev = min(max(prev, 0), total - 1)
    nxt = min(prev + 1, total - 1)
    return prev, nxt, float(vi - np.floor(vi))


def lerp(a, b, t):
    """numpy _lerp (numpy/lib/_function_base_impl.py:4657-4678)."""
    a, b, t = np.float64(a), np.float64(b), np.float64(t)
    diff = b - a
    if t >= 0.5:
        return float(b - diff * (1 - t))
    return float(a + diff * t)


def percentile_from_order_stats(total: int, q: float, value_of_rank) -> float:
    """np.percentile(sorted_values, q) given a callable rank -> value."""
    prev, nxt, gamma = percentile_ranks(total, q)
    return lerp(value_of_rank(prev), value_of_rank(nxt), gamma)


def median_from_order_stats(total: int, value_of_rank) -> float:
    """np.median: middle element, or the mean of the two middle ones
    (numpy/lib/_function_base_impl.py:4037-4073)."""
    if total % 2:
        return float(value_of_rank(total // 2))
    a, b = np.float64(value_of_rank(total // 2 - 1)), np.float64(value_of_rank(total // 2))
    return float((a + b) / 2.0)
