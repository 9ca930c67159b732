"""Oracle (test infrastructure, not product): the experimental part of `Variogram` /
`DirectionalVariogram` for point sets too large for the reference's O(N^2) arrays, on top of the
C sweeps of oracle/pairs_oracle.c.  Same sequence of steps as the reference constructor:

  maxlag  'median' -> np.median(distance)                       Variogram.py:1288-1289
  bins    even     -> np.linspace(0, maxlag, n + 1)[1:]         binning.py:10-40 (maxlag <= max(d) here)
          uniform  -> np.nanpercentile(d[mask & (d <= maxlag)]) binning.py:43-79
  bandwidth 'q33'  -> np.percentile(distance, 33)               DirectionalVariogram.py:503-509
  groups / estimator per lag class                              Variogram.py:1989-2092, estimators.py

tests/test_oracle_fullsize.py pins it to the dense (literal) oracle at a size both can run."""
from __future__ import annotations

import numpy as np

from . import pairs as P
from . import variogram as ov


def isotropic(coords, values, estimators, n_lags=25):
    """maxlag='median', even bins: dict(maxlag, bins, bin_count, <estimator>: experimental...)."""
    pr = P.Pairs(coords, values)
    vals, tot = pr.order_stats(lambda g, t: P.median_ranks(t))
    maxlag = P.median_from(tot[0], lambda r: vals[0][r])
    edges = np.linspace(0, maxlag, n_lags + 1)[1:]           # binning.py:40
    count, ssq, srt, _ = pr.lag_hist(edges)
    out = dict(maxlag=maxlag, bins=edges, bin_count=count, n_pairs=tot[0])
    for est in estimators:
        if est in ("matheron", "cressie"):
            out[est] = ov.finalize(count, ssq, srt, est)
        elif est == "dowd":
            mv, mt = pr.order_stats(lambda g, t: P.median_ranks(t), edges=edges, key_kind=1)
            assert [int(t) for t in mt] == [int(c) for c in count]
            med = np.array([P.median_from(mt[g], lambda r, g=g: mv[g][r]) if mt[g] else np.nan
                            for g in range(n_lags)])
            out["dowd_median"] = med
            out[est] = 2.198 * med ** 2 / 2.0                  # estimators.py:178
        else:
            raise ValueError(est)
    return out


def directional_uniform(coords, values, azimuth, tolerance, n_lags=10, bandwidth_q=33):
    """DirectionalVariogram(azimuth, tolerance, bin_func='uniform', maxlag=None, bandwidth='q33',
    directional_model='triangle', estimator='matheron')."""
    iso = P.Pairs(coords, None)
    ranks, _ = None, None
    vals, tot = iso.order_stats(lambda g, t: P.percentile_ranks(t, bandwidth_q)[0])
    bw = P.percentile_from(tot[0], bandwidth_q, lambda r: vals[0][r])          # :509
    dirp = (1, np.radians(azimuth), np.radians(tolerance / 2), bw / 2)
    pr = P.Pairs(coords, values, dirp)
    qs = [(i / n_lags) * 100 for i in range(1, n_lags + 1)]                     # binning.py:76

    def rank_fn(g, t):
        out = set()
        for q in qs:
            out.update(P.percentile_ranks(t, q)[0])
        return sorted(out)

    dv, dt = pr.order_stats(rank_fn)                  # masked distances (NaN-masked in the reference)
    edges = np.array([P.percentile_from(dt[0], q, lambda r: dv[0][r]) for q in qs])
    count, ssq, srt, masked = pr.lag_hist(edges)
    return dict(bandwidth=bw, bins=edges, bin_count=count, mask_count=dt[0], n_guard=pr.n_guard,
                matheron=ov.finalize(count, ssq, srt, "matheron"))
