"""Oracle (test infrastructure, not product): CPU restatement of the
experimental-variogram path of scikit-gstat 1.0.23.

Two forms are provided.

* ``dense_*``: literal restatement. Materialises the condensed distance and
  |dz| vectors with scipy ``pdist`` exactly like the reference does and applies
  numpy ``digitize`` / ``median`` / ``percentile``.  O(N^2) memory; use for
  N <= ~20k.  This is also the "port" timed as the CPU baseline.
* ``blocked_*``: row-blocked restatement that never holds more than
  ``block x N`` pairs.  Same per-pair arithmetic (verified bit-identical to
  scipy in tests/test_oracle_variogram.py), exact integer counts and
  ``math.fsum`` sums.  Use for N too large for the dense form, and as the
  high-precision checker where the reference's own sequential sums are noisy.

Every function cites the reference file:line it restates (paths relative to
``/root/reference/skgstat``).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.spatial.distance import pdist

# --------------------------------------------------------------------------
# dense (literal) form
# --------------------------------------------------------------------------


def coords_2d(coordinates) -> np.ndarray:
    """Variogram.py:286-297 - 1-D coordinates get a zero column appended."""
    c = np.asarray(coordinates)
    if c.ndim < 2:
        c = np.column_stack((c, np.zeros(len(c))))
    return c


def distance(coordinates) -> np.ndarray:
    """MetricSpace.py:151-153 + Variogram.py:1202-1208: condensed euclidean
    distances, order k = N*i - i(i+1)/2 + (j-i-1), i<j."""
    return pdist(coords_2d(coordinates), metric="euclidean")


def pairwise_diffs(values) -> np.ndarray:
    """Variogram.py:1938-1944: |z_i - z_j| through pdist on (z, 0)."""
    v = np.asarray(values)
    return pdist(np.column_stack((v, np.zeros(len(v)))), metric="euclidean")


def pairwise_diffs_multi(values, mode) -> np.ndarray:
    """Variogram.py:1965-1984 for 2-D values: mode 'cross' -> |dz_0| * |dz_1| (values[:, 0] is
    the observation, values[:, 1] the co-variable, :550-566); mode 'aggregate' ->
    sqrt(sum_k dz_k^2) with the rows stacked and summed along axis 0 (:1974-1977)."""
    v = np.asarray(values, dtype=float)
    if mode == "cross":
        return pairwise_diffs(v[:, 0]) * pairwise_diffs(v[:, 1])
    if mode == "aggregate":
        diffs = np.vstack([pairwise_diffs(v[:, k]) for k in range(v.shape[1])])
        return np.sqrt(np.sum(diffs ** 2, axis=0))
    raise ValueError(mode)


def resolve_maxlag(d: np.ndarray, maxlag):
    """Variogram.py:1274-1295 (maxlag.setter)."""
    if maxlag is None:
        return None
    if isinstance(maxlag, str):
        if maxlag == "median":
            return np.median(d)
        if maxlag == "mean":
            return np.mean(d)
        return None  # the reference silently leaves _maxlag = None
    if maxlag < 1:
        return maxlag * np.max(d)
    return maxlag


def even_width_lags(d: np.ndarray, n: int, maxlag):
    """binning.py:10-40."""
    if maxlag is None or maxlag > np.nanmax(d):
        maxlag = np.nanmax(d)
    return np.linspace(0, maxlag, n + 1)[1:]


def uniform_count_lags(d: np.ndarray, n: int, maxlag):
    """binning.py:43-79."""
    if maxlag is None or maxlag > np.nanmax(d):
        maxlag = np.nanmax(d)
    dd = d[np.where(d <= maxlag)]
    return np.fromiter(
        (np.nanpercentile(dd, (i / n) * 100) for i in range(1, n + 1)), dtype=float
    )


def auto_derived_lags(d: np.ndarray, method_name: str, maxlag):
    """binning.py:82-122: np.histogram_bin_edges rules ('sqrt', 'sturges', 'rice', 'scott',
    'doane', 'fd', 'auto') on the distances up to maxlag."""
    if maxlag is None or maxlag > np.nanmax(d):
        maxlag = np.nanmax(d)
    dd = d[np.where(d <= maxlag)]
    return np.histogram_bin_edges(dd, bins=method_name)[1:]


AUTO_BINNERS = ("sqrt", "sturges", "rice", "scott", "doane", "fd", "auto")
BIN_FUNCS = {"even": even_width_lags, "uniform": uniform_count_lags}
BIN_FUNCS.update({m: (lambda d, n, ml, m=m: auto_derived_lags(d, m, ml)) for m in AUTO_BINNERS})


def lag_groups(d: np.ndarray, edges: np.ndarray) -> np.ndarray:
    """Variogram.py:1989-2006: digitize, index n_lags -> -1."""
    g = np.digitize(d, edges)
    return np.where(g == len(edges), -1, g)


def matheron(x: np.ndarray) -> float:
    """estimators.py:63-66."""
    if x.size == 0:
        return np.nan
    return (1.0 / (2 * x.size)) * np.sum(np.power(x, 2))


def cressie(x: np.ndarray) -> float:
    """estimators.py:112-128 (n**2 evaluated in float to avoid the int64
    overflow the numba build has for n > 3.03e9)."""
    n = x.size
    if n == 0:
        return np.nan
    nominator = np.power((1 / n) * np.sum(np.power(x, 0.5)), 4)
    denominator = 0.457 + (0.494 / n) + (0.045 / float(n) ** 2)
    return nominator / (2 * denominator)


def dowd(x: np.ndarray) -> float:
    """estimators.py:178."""
    if x.size == 0:
        return np.nan
    return 2.198 * np.nanmedian(x) ** 2 / 2


def minmax(x: np.ndarray) -> float:
    """estimators.py:296."""
    return (np.nanmax(x) - np.nanmin(x)) / np.nanmean(x)


def percentile(x: np.ndarray, p=50) -> float:
    """estimators.py:324."""
    return np.percentile(x, q=p)


def shannon_entropy(x: np.ndarray, bins) -> float:
    """util/shannon.py:27-34 (estimators.entropy, estimators.py:327-367)."""
    c, _ = np.histogram(x, bins=bins)
    with np.errstate(divide="ignore", invalid="ignore"):
        p = c / np.sum(c) + 1e-15
    return -np.fromiter(map(np.log2, p), dtype=float).dot(p)


ESTIMATORS = {"matheron": matheron, "cressie": cressie, "dowd": dowd, "minmax": minmax,
              "percentile": percentile}


def dense_variogram(coordinates, values, estimator="matheron", bin_func="even",
                    n_lags=10, maxlag=None, multivariate=None, max_dist=None, **kwargs):
    """Variogram.py:31-405 (constructor) restricted to the experimental part:
    returns dict(maxlag, bins, bin_count, experimental).  multivariate: None (1-D values),
    'cross' or 'aggregate' (2-D values).  max_dist: sparse MetricSpace (MetricSpace.py:141-147,
    Variogram.py:1210-1227) - only the pairs with 0 < d <= max_dist exist.  kwargs: 'percentile'
    and 'entropy_bins' as Variogram.py:2065-2082 reads them."""
    d = distance(coordinates)
    if multivariate is None:
        diffs = pairwise_diffs(np.asarray(values, dtype=float))
    else:
        diffs = pairwise_diffs_multi(values, multivariate)
    if max_dist is not None:
        # the reference's sparse vectors hold the lower-triangle entries row by row
        # (Variogram.py:1210-1227), i.e. the pairs (i < j) ordered by j, then i - the order
        # matters for np.mean's pairwise summation (maxlag='mean')
        keep = np.flatnonzero((d > 0) & (d <= max_dist))
        iu, ju = np.triu_indices(len(np.asarray(coordinates)), 1)
        keep = keep[np.lexsort((iu[keep], ju[keep]))]
        d, diffs = d[keep], diffs[keep]
    ml = resolve_maxlag(d, maxlag)
    edges = BIN_FUNCS[bin_func](d, n_lags, ml)
    g = lag_groups(d, edges)
    if estimator == "entropy":          # Variogram.py:2065-2075: edges from the DISTANCE vector
        nb = kwargs.get("entropy_bins", 50)
        if isinstance(nb, int):
            nb -= 1
        vbins = np.histogram_bin_edges(d, bins=nb)
        est = lambda x: shannon_entropy(x, vbins)
    elif estimator == "percentile" and kwargs.get("percentile", False):
        est = lambda x: percentile(x, kwargs["percentile"])
    else:
        est = ESTIMATORS[estimator]
    # Variogram.py:1556-1557 / 2092: one boolean sweep per lag class
    classes = [diffs[np.where(g == i)] for i in range(len(edges))]
    return dict(
        maxlag=ml,
        bins=edges,
        bin_count=np.fromiter((c.size for c in classes), dtype=np.int64),
        experimental=np.fromiter((est(c) for c in classes), dtype=float),
    )


# ---- directional ---------------------------------------------------------


def direction_angles(coordinates):
    """DirectionalVariogram.py:348-413, with the two pdist(..., np.subtract)
    Python-callback passes (:404, :409) replaced by the vectorised
    x_i - x_j (bit-identical: a single IEEE subtraction either way).
    Returns (euclidean_dist, angles), both condensed."""
    x = np.asarray(coordinates, dtype=float)
    n = len(x)
    i, j = np.triu_indices(n, k=1)
    ed = pdist(x, "euclidean")
    scalar = x[i, 0] - x[j, 0]
    with np.errstate(invalid="ignore", divide="ignore"):
        pos = np.arccos(scalar / ed)
    ydiff = x[i, 1] - x[j, 1]
    return ed, np.where(ydiff >= 0, pos, -pos)


def triangle_mask(angles, dists, azimuth, tolerance, bandwidth):
    """DirectionalVariogram.py:659-714."""
    with np.errstate(invalid="ignore"):
        absdiff = np.abs(angles + np.radians(azimuth))
        absdiff = np.where(absdiff > np.pi, absdiff - np.pi, absdiff)
        absdiff = np.where(absdiff > np.pi / 2, np.pi - absdiff, absdiff)
        in_tol = absdiff <= np.radians(tolerance / 2)
        in_band = bandwidth / 2 >= np.abs(
            dists * np.sin(np.abs(angles + np.radians(azimuth))))
    return in_tol & in_band


def compass_mask(angles, dists, azimuth, tolerance, bandwidth=None):
    """DirectionalVariogram.py:749-782."""
    with np.errstate(invalid="ignore"):
        absdiff = np.abs(angles + np.radians(azimuth))
        absdiff = np.where(absdiff > np.pi, absdiff - np.pi, absdiff)
        absdiff = np.where(absdiff > np.pi / 2, np.pi - absdiff, absdiff)
        return absdiff <= np.radians(tolerance / 2)


def resolve_bandwidth(d, bandwidth):
    """DirectionalVariogram.py:497-516 ('q33' -> np.percentile(d, 33)).  A width beyond the largest
    distance is NOT stored (:513-515 only print a note), so the property falls back to 0 (:498-499)
    and the triangle's strip collapses onto the azimuth axis - restated as is."""
    if isinstance(bandwidth, str):
        return np.percentile(d, int(bandwidth[1:]))
    if bandwidth > np.max(d):
        return 0
    return bandwidth


def dense_directional(coordinates, values, estimator="matheron", bin_func="even",
                      n_lags=10, maxlag=None, azimuth=0, tolerance=45.0,
                      bandwidth="q33", directional_model="triangle"):
    """DirectionalVariogram.py:23-346, 577-624: experimental part."""
    c = np.asarray(coordinates, dtype=float)
    d = distance(c)
    diffs = pairwise_diffs(np.asarray(values, dtype=float))
    ml = resolve_maxlag(d, maxlag)
    bw = resolve_bandwidth(d, bandwidth)
    ed, ang = direction_angles(c)
    if directional_model == "triangle":
        mask = triangle_mask(ang, ed, azimuth, tolerance, bw)
    else:
        mask = compass_mask(ang, ed, azimuth, tolerance)
    dm = d.copy()
    dm[np.where(~mask)] = np.nan
    with np.errstate(invalid="ignore"):
        edges = BIN_FUNCS[bin_func](dm, n_lags, ml)
    g = lag_groups(d, edges)
    g[np.where(~mask)] = -1
    est = ESTIMATORS[estimator]
    classes = [diffs[np.where(g == i)] for i in range(len(edges))]
    return dict(
        maxlag=ml, bandwidth=bw, bins=edges, mask_count=int(mask.sum()),
        bin_count=np.fromiter((c_.size for c_ in classes), dtype=np.int64),
        experimental=np.fromiter((est(c_) for c_ in classes), dtype=float),
    )


# --------------------------------------------------------------------------
# blocked form (no O(N^2) storage)
# --------------------------------------------------------------------------


def _row_block(c: np.ndarray, z: np.ndarray, i0: int, i1: int):
    """All pairs (i, j), i0 <= i < i1, j > i: returns d, |dz|, dx, dy (1-D,
    in condensed order).  d = sqrt(fl(fl(dx*dx)+fl(dy*dy)) [+ fl(dz*dz)])
    accumulated left-to-right over the dimensions without FMA - the operation
    order of scipy's euclidean kernel."""
    n, dim = c.shape
    out_d, out_z, out_dx, out_dy = [], [], [], []
    for i in range(i0, i1):
        diff = c[i] - c[i + 1:]
        s = diff[:, 0] * diff[:, 0]
        for k in range(1, dim):
            s = s + diff[:, k] * diff[:, k]
        out_d.append(np.sqrt(s))
        out_z.append(np.abs(z[i] - z[i + 1:]))
        out_dx.append(diff[:, 0])
        out_dy.append(diff[:, 1] if dim > 1 else np.zeros(n - i - 1))
    cat = np.concatenate
    return cat(out_d), cat(out_z), cat(out_dx), cat(out_dy)


def blocked_angles(dx, dy, d):
    """Same arithmetic as direction_angles for one block."""
    with np.errstate(invalid="ignore", divide="ignore"):
        pos = np.arccos(dx / d)
    return np.where(dy >= 0, pos, -pos)


def blocked_hist(coordinates, values, edges, block=256, mask=None,
                 want_median=False):
    """Exact per-bin count, fsum(dz^2), fsum(sqrt|dz|) for given upper edges.

    mask: None or dict(model, azimuth, tolerance, bandwidth(float)).
    Returns dict(count[int64], sum_sq[float], sum_sqrt[float], masked[int],
    classes (list of |dz| arrays, only if want_median)).
    """
    c = coords_2d(np.asarray(coordinates, dtype=float))
    z = np.asarray(values, dtype=float)
    n = len(c)
    nl = len(edges)
    count = np.zeros(nl, dtype=np.int64)
    part_sq = [[] for _ in range(nl)]
    part_rt = [[] for _ in range(nl)]
    classes = [[] for _ in range(nl)]
    masked = 0
    for i0 in range(0, n - 1, block):
        i1 = min(n - 1, i0 + block)
        d, dz, dx, dy = _row_block(c, z, i0, i1)
        g = lag_groups(d, edges)
        if mask is not None:
            ang = blocked_angles(dx, dy, d)
            fn = triangle_mask if mask["model"] == "triangle" else compass_mask
            m = fn(ang, d, mask["azimuth"], mask["tolerance"], mask["bandwidth"])
            masked += int(m.sum())
            g[~m] = -1
        for k in range(nl):
            x = dz[g == k]
            count[k] += x.size
            if x.size:
                part_sq[k].append(math.fsum(x * x))
                part_rt[k].append(math.fsum(np.sqrt(x)))
                if want_median:
                    classes[k].append(x)
    out = dict(
        count=count,
        sum_sq=np.array([math.fsum(p) for p in part_sq]),
        sum_sqrt=np.array([math.fsum(p) for p in part_rt]),
        masked=masked,
    )
    if want_median:
        out["classes"] = [np.concatenate(k) if k else np.zeros(0) for k in classes]
    return out


def finalize(count, sum_sq, sum_sqrt, estimator):
    """estimators.py:66 / :123-128 applied to accumulated sums."""
    n = np.asarray(count, dtype=float)
    with np.errstate(invalid="ignore", divide="ignore"):
        if estimator == "matheron":
            out = (1.0 / (2 * n)) * np.asarray(sum_sq)
        elif estimator == "cressie":
            nom = np.power((1 / n) * np.asarray(sum_sqrt), 4)
            den = 0.457 + (0.494 / n) + (0.045 / n ** 2)
            out = nom / (2 * den)
        else:
            raise ValueError(estimator)
    out = np.where(np.asarray(count) == 0, np.nan, out)
    return out
