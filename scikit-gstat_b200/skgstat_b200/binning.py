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


def auto_derived_lags(distances, method_name, maxlag):
    """binning.py:82-122 (operator form over a materialised vector)."""
    d = np.asarray(distances, dtype=float)
    if maxlag is None or maxlag > np.nanmax(d):
        maxlag = np.nanmax(d)
    d = d[np.where(d <= maxlag)]
    edges = np.histogram_bin_edges(d, bins=method_name)[1:]
    return edges, len(edges)


AUTO_BINNERS = ("sqrt", "sturges", "rice", "scott", "doane", "fd", "auto")

# ---- scalar forms used with the device-delivered statistics -----------------

def auto_bin_width(method, n, dmin, dmax, q25=None, q75=None, std=None, g1=None):
    """numpy's bin-width rules (numpy/lib/_histograms_impl.py: _hist_bin_sqrt ... _hist_bin_auto,
    called by np.histogram_bin_edges in binning.py:120) evaluated on scalars: n values with
    minimum dmin, maximum dmax, quartiles, standard deviation and skewness g1."""
    if method not in AUTO_BINNERS:
        raise ValueError("%r is not a valid estimator for `bins`" % method)
    ptp = np.float64(dmax) - np.float64(dmin)
    if method == "sqrt":
        return ptp / np.sqrt(n)
    if method == "sturges":
        return ptp / (np.log2(n) + 1.0)
    if method == "rice":
        return ptp / (2.0 * n ** (1.0 / 3))
    if method == "scott":
        return (24.0 * np.pi ** 0.5 / n) ** (1.0 / 3.0) * std
    if method == "doane":
        if n > 2:
            sg1 = np.sqrt(6.0 * (n - 2) / ((n + 1.0) * (n + 3)))
            if std > 0.0:
                return ptp / (1.0 + np.log2(n) + np.log2(1.0 + np.absolute(g1) / sg1))
        return 0.0
    fd = 2.0 * np.subtract(q75, q25) * n ** (-1.0 / 3.0)
    if method == "fd":
        return fd
    sturges = ptp / (np.log2(n) + 1.0)     # "auto"
    return min(max(fd, (ptp / np.sqrt(n)) / 2), sturges)


def auto_edges(method, n, dmin, dmax, **stats):
    """np.histogram_bin_edges(d, bins=method)[1:] from scalars (_get_bin_edges:
    n_bins = ceil(ptp / width), edges = linspace(min, max, n_bins + 1))."""
    first, last = np.float64(dmin), np.float64(dmax)
    if first == last:                      # _get_outer_edges
        first, last = first - 0.5, last + 0.5
    if n == 0:
        n_bins = 1
    else:
        width = auto_bin_width(method, n, dmin, dmax, **stats)
        n_bins = int(np.ceil((last - first) / width)) if width else 1
    edges = np.linspace(first, last, n_bins + 1, endpoint=True, dtype=np.float64)
    if np.any(edges[:-1] >= edges[1:]):
        raise ValueError("Too many bins for data range. Cannot create %d finite-sized bins." % n_bins)
    return edges[1:]


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
