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
    # co-located samples: the reference removes the point and THEN de-duplicates
    # (util/cross_validation.py:9-21 -> OrdinaryKriging._remove_duplicated_coordinates), so the
    # target keeps a neighbour at distance zero.  Those (rare) targets are kriged one at a time on
    # the sample set without the point, exactly as upstream; all others share one batched call.
    _, inverse, counts = np.unique(c, axis=0, return_inverse=True, return_counts=True)
    dup_point = counts[np.ravel(inverse)] > 1
    out = np.full(len(indices), np.nan)
    plain = ~dup_point[indices]
    if plain.any():
        if dup_point.any():
            # with twins present the batch runs on the de-duplicated set (first occurrences, as the
            # reference keeps) and excludes each target by its position in that set
            first = np.sort(np.unique(c, axis=0, return_index=True)[1])
            pos = -np.ones(len(c), dtype=np.int64)
            pos[first] = np.arange(first.size)
            cs, vs = c[first], v[first]
        else:
            pos = np.arange(len(c))
            cs, vs = c, v
        ok = OrdinaryKriging(variogram, min_points=min_points, max_points=max_points,
                             coordinates=cs, values=vs)
        eng = ok._engine()
        idx = indices[plain]
        z, sigma, status = eng.transform(c[idx], ok._model_name, ok._model_params(),
                                         ok._minp, ok._maxp, exclude=pos[idx].astype(np.int32))
        out[plain] = z - v[idx].astype(np.float64)
    for k in np.flatnonzero(~plain):
        i = int(indices[k])
        ok = OrdinaryKriging(variogram, min_points=min_points, max_points=max_points,
                             coordinates=np.delete(c, i, axis=0), values=np.delete(v, i))
        z = ok.transform(*[np.atleast_1d(x) for x in c[i]])
        out[k] = float(z[0]) - float(v[i])
    return out


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
