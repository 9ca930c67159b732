"""Semi-variance estimators.

On the hot path the per-lag reductions (count, sum dz^2, sum sqrt|dz|, median
|dz|) come from the fused pair kernels; ``finalize_*`` turn them into
semi-variances with the reference's formulas.  The array forms
(``matheron(x)`` ...) keep the reference's operator signature
``estimator(x) -> float`` (estimators.py:14-178) for callers that already hold
a materialised lag class.
"""
from __future__ import annotations

import numpy as np


def matheron(x):
    """estimators.py:63-66."""
    x = np.asarray(x, dtype=float)
    if x.size == 0:
        return np.nan
    return (1.0 / (2 * x.size)) * np.sum(np.power(x, 2))


def cressie(x):
    """estimators.py:112-128."""
    x = np.asarray(x, dtype=float)
    n = x.size
    if n == 0:
        return np.nan
    nominator = np.power((1 / n) * np.sum(np.power(x, 0.5)), 4)
    denominator = 0.457 + (0.494 / n) + (0.045 / float(n) ** 2)
    return nominator / (2 * denominator)


def dowd(x):
    """estimators.py:178."""
    x = np.asarray(x, dtype=float)
    if x.size == 0:
        return np.nan
    return 2.198 * np.nanmedian(x) ** 2 / 2


def finalize_matheron(count, sum_sq):
    n = np.asarray(count, dtype=float)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = (1.0 / (2 * n)) * np.asarray(sum_sq, dtype=float)
    return np.where(n == 0, np.nan, out)


def finalize_cressie(count, sum_sqrt):
    n = np.asarray(count, dtype=float)
    with np.errstate(divide="ignore", invalid="ignore"):
        nominator = np.power((1 / n) * np.asarray(sum_sqrt, dtype=float), 4)
        denominator = 0.457 + (0.494 / n) + (0.045 / n ** 2)
        out = nominator / (2 * denominator)
    return np.where(n == 0, np.nan, out)


def finalize_dowd(count, median):
    m = np.asarray(median, dtype=float)
    return np.where(np.asarray(count) == 0, np.nan, 2.198 * m ** 2 / 2)


def minmax(x):
    """estimators.py:257-296."""
    x = np.asarray(x, dtype=float)
    return (np.nanmax(x) - np.nanmin(x)) / np.nanmean(x)


def percentile(x, p=50):
    """estimators.py:299-324."""
    return np.percentile(np.asarray(x, dtype=float), q=p)


def entropy(x, bins=None):
    """estimators.py:327-367 + util/shannon.py:27-34."""
    if bins is None:
        bins = 15
    c, _ = np.histogram(np.asarray(x, dtype=float), bins=bins)
    return shannon_from_counts(c)


def shannon_from_counts(c):
    """util/shannon.py:30-34 on a histogram that already exists."""
    c = np.asarray(c)
    with np.errstate(divide="ignore", invalid="ignore"):
        p = c / np.sum(c) + 1e-15
    return -np.fromiter(map(np.log2, p), dtype=float).dot(p)


def finalize_minmax(count, vmin, vmax, total):
    n = np.asarray(count, dtype=float)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = (np.asarray(vmax, dtype=float) - np.asarray(vmin, dtype=float)) / (np.asarray(total, dtype=float) / n)
    return np.where(n == 0, np.nan, out)


ESTIMATORS = dict(matheron=matheron, cressie=cressie, dowd=dowd, minmax=minmax,
                  percentile=percentile, entropy=entropy)
