"""Cross-variogram grid (reference: skgstat/util/cross_variogram.py:14-87).

For (n_samples, N) values an N x N nested list of variograms is returned: primary variograms
on the diagonal, cross-variograms (|dz_i| * |dz_j|) elsewhere.  All N^2 instances share ONE
MetricSpace, i.e. one device-resident, Z-order sorted copy of the coordinates, and the bins are
derived once per instance from exact device order statistics - the reference recomputes the
full distance matrix for every cell.
"""
from __future__ import annotations

import numpy as np

from ..DirectionalVariogram import DirectionalVariogram
from ..MetricSpace import MetricSpace
from ..Variogram import Variogram


def cross_variograms(coordinates, values, **kwargs):
    """skgstat/util/cross_variogram.py:14-87 (same arguments, same return structure)."""
    values = np.asarray(values)
    if any(arg in kwargs for arg in ("azimuth", "tolerance", "bandwidth")):
        base_cls = DirectionalVariogram
    else:
        base_cls = Variogram
    if not isinstance(coordinates, MetricSpace):
        c = np.asarray(coordinates)
        if c.ndim >= 2:  # 1-D input keeps the reference's zero-column handling in Variogram
            coordinates = MetricSpace(c.copy(), kwargs.get("dist_func", "euclidean"), None)
    n_var = values.shape[1]
    cross_m = []
    for i in range(n_var):
        row = []
        for j in range(n_var):
            if i == j:
                row.append(base_cls(coordinates, values[:, i], **kwargs))
            else:
                row.append(base_cls(coordinates, values[:, [i, j]], **kwargs))
        cross_m.append(row)
    return cross_m
