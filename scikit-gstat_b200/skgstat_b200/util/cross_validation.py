"""Leave-one-out ("jacknife") cross validation on the batched kriging kernel.

The reference (util/cross_validation.py:9-74) loops over the sampled points in Python:
for each index it deletes the point, builds a fresh ``OrdinaryKriging`` (N x N pdist of
the remaining samples) and krigs at the deleted location.  Kriging only ever sees the
neighbours of the target, so the same numbers come from ONE ``skg_krige`` call whose
targets are the sample locations and whose ``exclude[t]`` is the sample itself.
"""
from __future__ import annotations

import numpy as np

from ..Kriging import OrdinaryKriging


def leave_one_out(variogram, indices=None, min_points=5, max_points=15, coordinates=None,
                  values=None):
    """Deviations Z*(x_i) - Z(x_i) for the given sample indices (all by default).
    `variogram` is a Variogram or a describe()-style dict (then pass coordinates/values)."""
    c = np.asarray(variogram.coordinates if coordinates is None else coordinates, dtype=np.float64)
    v = np.asarray(variogram.values if values is None else values)
    if indices is None:
        indices = np.arange(len(c))
    indices = np.asarray(indices, dtype=np.int64)
    if len(np.unique(c, axis=0)) != len(c):
        raise NotImplementedError("jacknife with duplicated sample coordinates is not built "
                                  "(the reference de-duplicates after removing the point)")
    ok = OrdinaryKriging(variogram, min_points=min_points, max_points=max_points,
                         coordinates=c, values=v)
    eng = ok._engine()
    z, sigma, status = eng.transform(c[indices], ok._model_name, ok._model_params(),
                                     ok._minp, ok._maxp, exclude=indices.astype(np.int32))
    return z - v[indices].astype(np.float64)


def jacknife(variogram, n=None, metric="rmse", seed=None):
    """util/cross_validation.py:24-74: same sampling (default_rng(seed).choice), same metrics."""
    if metric.lower() not in ("rmse", "mse", "mae"):
        raise ValueError("metric has to be in ['rmse', 'mse', 'mae']")
    rng = np.random.default_rng(seed=seed)
    size = n if n is not None else len(variogram.coordinates)
    indices = rng.choice(len(variogram.coordinates), replace=False, size=size)
    deviations = leave_one_out(variogram, indices)
    if metric.lower() == "rmse":
        return np.sqrt(np.nanmean(np.power(deviations, 2)))
    if metric.lower() == "mse":
        return np.nanmean(np.power(deviations, 2))
    return np.nansum(np.abs(deviations)) / len(deviations)
