"""Oracle (test infrastructure, not product): CPU restatement of
``OrdinaryKriging.transform`` of scikit-gstat 1.0.23 (mode='exact',
solver='inv', n_jobs=1 - the reference defaults; the n_jobs>1 branch of the
reference returns NaN for every point and is not restated).

Paths relative to ``/root/reference/skgstat``.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.distance import cdist, pdist, squareform


# ---- theoretical models, models.py:21-344 (vectorised; same formulas) -----

def spherical(h, r, c0, b=0.0):
    """models.py:77-82."""
    h = np.asarray(h, dtype=float)
    a = r / 1.0
    inside = b + c0 * ((1.5 * (h / a)) - (0.5 * ((h / a) ** 3.0)))
    return np.where(h <= r, inside, b + c0)


def exponential(h, r, c0, b=0.0):
    """models.py:142-144."""
    h = np.asarray(h, dtype=float)
    a = r / 3.0
    return b + c0 * (1.0 - np.exp(-(h / a)))


def gaussian(h, r, c0, b=0.0):
    """models.py:207-209."""
    h = np.asarray(h, dtype=float)
    a = r / 2.0
    return b + c0 * (1.0 - np.exp(-(h ** 2 / a ** 2)))


def cubic(h, r, c0, b=0.0):
    """models.py:264-272 (note the strict h < r)."""
    h = np.asarray(h, dtype=float)
    a = r / 1.0
    inside = b + c0 * ((7 * (h ** 2 / a ** 2)) - ((35 / 4) * (h ** 3 / a ** 3))
                       + ((7 / 2) * (h ** 5 / a ** 5)) - ((3 / 4) * (h ** 7 / a ** 7)))
    return np.where(h < r, inside, b + c0)


def stable(h, r, c0, s, b=0.0):
    """models.py:336-344 (h == 0 -> b)."""
    h = np.asarray(h, dtype=float)
    a = r / np.power(3, 1 / s)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = b + c0 * (1.0 - np.exp(-np.power(h / a, s)))
    return np.where(h == 0, b, v)


def matern(h, r, c0, s, b=0.0):
    """models.py:413-422 (h == 0 -> b; scipy.special.kv / gamma)."""
    from scipy import special
    h = np.asarray(h, dtype=float)
    a = r / 2.0
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        v = b + c0 * (1.0 - (2 / special.gamma(s)) * np.power((h * np.sqrt(s)) / a, s)
                      * special.kv(s, 2 * ((h * np.sqrt(s)) / a)))
    return np.where(h == 0, b, v)


MODELS = dict(spherical=spherical, exponential=exponential, gaussian=gaussian,
              cubic=cubic, stable=stable, matern=matern)


def fitted_model(model, effective_range, sill, nugget=0.0, shape=None):
    """Variogram.py:1870-1900 (fitted_model_function): cof = [range, sill,
    (shape), (nugget if != 0)]."""
    fn = MODELS[model]
    cof = [effective_range, sill]
    if shape is not None:
        cof.append(shape)
    if nugget != 0:
        cof.append(nugget)
    return lambda h: fn(h, *cof)


def remove_duplicated_coordinates(coords, values):
    """Kriging.py:285-303."""
    _, idx = np.unique(coords, axis=0, return_index=True)
    idx.sort()
    return coords[idx], values[idx]


def find_closest(dist_row, max_dist, n):
    """MetricSpace.py:26-71 (dense branch)."""
    ridx = np.where(dist_row <= max_dist)[0]
    if ridx.size > n:
        order = np.argsort(dist_row[ridx], kind="stable")
        ridx = ridx[order][:n]
    return ridx


def estimate_matrix(dist, gamma, effective_range, sill, precision):
    """Kriging.py:656-679 (mode='estimate'): semivariances of the kriging matrix from a table
    of `precision` values over [0, range]; beyond the table the sill."""
    prec_g = gamma(np.linspace(0, effective_range, precision))
    dist_n = ((dist / effective_range) * precision).astype(int)
    g = np.ones(dist_n.shape) * -1
    out_ = np.where(dist_n >= precision)[0]
    in_ = np.where(dist_n < precision)[0]
    g[out_] = sill
    g[in_] = prec_g[dist_n[in_]]
    return g


def krige(coords, values, targets, model, effective_range, sill, nugget=0.0,
          shape=None, min_points=5, max_points=15, chunk=2048, mode="exact", precision=100):
    """Kriging.py:405-650.  Returns (z, sigma, status) with status 0 = ok,
    1 = fewer than min_points neighbours (LessPointsError -> NaN),
    2 = singular matrix (numpy LinAlgError; the reference would let that
    propagate, Kriging.py:618-623 - the oracle records NaN instead)."""
    coords = np.asarray(coords, dtype=float)
    values = np.asarray(values)
    coords, values = remove_duplicated_coordinates(coords, values)
    targets = np.asarray(targets, dtype=float)
    gamma = fitted_model(model, effective_range, sill, nugget, shape)
    m = len(targets)
    z = np.full(m, np.nan)
    sigma = np.full(m, np.nan)
    status = np.zeros(m, dtype=np.int32)
    for c0 in range(0, m, chunk):
        # MetricSpace.py:248-253: cdist(targets, samples)
        dmat = cdist(targets[c0:c0 + chunk], coords, metric="euclidean")
        for r in range(dmat.shape[0]):
            t = c0 + r
            idx = find_closest(dmat[r], effective_range, max_points)
            if idx.size < min_points:
                status[t] = 1
                continue
            in_range = coords[idx]
            vals = values[idx]
            # MetricSpace.py:158-187: sub-matrix of the cached pdist matrix
            dist_mat = pdist(in_range, metric="euclidean")
            n = len(in_range)
            if mode == "exact":                              # Kriging.py:594-600
                a = squareform(gamma(dist_mat))
            else:
                a = squareform(estimate_matrix(dist_mat, gamma, effective_range, sill, precision))
            a = np.concatenate((a, np.ones((n, 1))), axis=1)
            a = np.concatenate((a, np.ones((1, n + 1))), axis=0)
            a[-1, -1] = 0
            _p = np.concatenate(([targets[t]], in_range))
            _d = pdist(_p, metric="euclidean")[: n]          # Kriging.py:611-612
            _g = gamma(_d)
            b = np.concatenate((_g, [1]))
            try:
                lam = np.linalg.inv(a).dot(b)                # Kriging.py:137-138
            except np.linalg.LinAlgError:
                status[t] = 2
                continue
            sigma[t] = sum(b[:-1] * lam[:-1]) + lam[-1]      # Kriging.py:644
            z[t] = lam[:-1].dot(vals)                        # Kriging.py:647
    return z, sigma, status


def jacknife_deviations(coords, values, indices, model, effective_range, sill, nugget=0.0,
                        shape=None, min_points=5, max_points=15):
    """util/cross_validation.py:9-21 (_interpolate) for every index: delete the sample,
    krige at its location from the remaining ones, return Z* - Z."""
    coords = np.asarray(coords, dtype=float)
    values = np.asarray(values)
    out = np.empty(len(indices))
    for k, idx in enumerate(indices):
        c = np.delete(coords, idx, axis=0)
        v = np.delete(values, idx, axis=0)
        z, _, _ = krige(c, v, coords[idx][None, :], model, effective_range, sill, nugget, shape,
                        min_points, max_points)
        out[k] = z[0] - values[idx]
    return out
