"""GPU: the register-window fused kernel, the bulk-counting count sweep and the windowed selects on
inputs built to stress their branches at sizes where they are active: lattice coordinates (pairs
exactly ON lag edges), duplicated points, NaN coordinates / NaN observations, 1-D coordinates
(degenerate boxes), ragged sizes, few wide lag classes (1-, 2- and 3-class windows) - against the
row-blocked oracle with exact integer counts."""
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import variogram as ov

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def skg():
    import skgstat_b200
    skgstat_b200._lib.load()
    return skgstat_b200


def _blocked_reference(c, v, edges):
    """count, sum dz^2, sum sqrt|dz| per lag class by row blocks (oracle._row_block arithmetic)."""
    c2 = ov.coords_2d(np.asarray(c, dtype=float))
    z = np.asarray(v, dtype=float)
    n, nl = len(c2), len(edges)
    cnt = np.zeros(nl, dtype=np.int64)
    ssq = np.zeros(nl)
    srt = np.zeros(nl)
    for i0 in range(0, n - 1, 256):
        i1 = min(n - 1, i0 + 256)
        d, dz, _, _ = ov._row_block(c2, z, i0, i1)
        with np.errstate(invalid="ignore"):
            g = ov.lag_groups(d, edges)
        ok = g >= 0
        cnt += np.bincount(g[ok], minlength=nl)
        fin = ok & np.isfinite(dz)
        ssq += np.bincount(g[fin], weights=dz[fin] ** 2, minlength=nl)
        srt += np.bincount(g[fin], weights=np.sqrt(dz[fin]), minlength=nl)
    return cnt, ssq, srt


CASES = [
    # name, n, generator kind, n_lags
    ("uniform-2lags", 9000, "uniform", 2),
    ("uniform-5lags", 12000, "uniform", 5),
    ("lattice-ties", 8100, "lattice", 6),
    ("duplicates", 7000, "dup", 4),
    ("one-dimensional", 6000, "line", 3),
    ("nan-coordinates", 7001, "nan", 4),
    ("clustered-25lags", 15000, "clustered", 25),
    ("ragged-513", 513 * 9 + 7, "uniform", 3),
]


def _make(kind, n, rng):
    if kind == "lattice":
        side = int(np.sqrt(n))
        gx, gy = np.meshgrid(np.arange(side, dtype=float), np.arange(side, dtype=float))
        c = np.column_stack((gx.ravel(), gy.ravel()))[:n] * 3.0
    elif kind == "line":
        c = rng.random(n) * 5000.0
    elif kind == "clustered":
        c = rng.normal(0, 30, size=(n, 2)) + rng.integers(0, 6, size=(n, 2)) * 400.0
    else:
        c = rng.random((n, 2)) * 1000.0
    if kind == "dup":
        c[rng.integers(0, n, size=n // 5)] = c[rng.integers(0, 50, size=n // 5)]
    if kind == "nan":
        c[rng.integers(0, n, size=25), rng.integers(0, 2, size=25)] = np.nan
    v = rng.normal(size=n) * 3.0 + 10.0
    if kind == "lattice":
        v = np.round(v)        # tied observations too: |dz| = 0 pairs for the sqrt path
    return c, v


@pytest.mark.parametrize("name,n,kind,n_lags", CASES)
def test_window_paths_match_blocked_oracle(skg, name, n, kind, n_lags):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 32))
    c, v = _make(kind, n, rng)
    c2 = ov.coords_2d(np.asarray(c, dtype=float))
    ext = np.nanmax(c2, axis=0) - np.nanmin(c2, axis=0)
    maxlag = 0.45 * float(np.hypot(*ext))
    if kind == "lattice":
        maxlag = 3.0 * 30      # an edge sequence that lands exactly on lattice distances
    edges = np.linspace(0, maxlag, n_lags + 1)[1:]
    cnt, ssq, srt = _blocked_reference(c, v, edges)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for est, ref in (("matheron", ssq), ("cressie", srt)):
            V = skg.Variogram(c, v, estimator=est, n_lags=n_lags, maxlag=maxlag, fit_method=None)
            assert_array_equal(V.bins, edges)
            assert_array_equal(V.bin_count, cnt, err_msg="%s %s" % (name, est))
            eng = V._engine()
            from skgstat_b200 import _lib
            got = eng.hist(edges, _lib.STAT_SUM_SQ if est == "matheron" else _lib.STAT_SUM_SQRT)
            assert_array_equal(got[0], cnt)
            assert_allclose(got[1] if est == "matheron" else got[2], ref, rtol=1e-12, atol=1e-300)
        # count-only sweep (bulk + windows) and the windowed Dowd sweep against the same counts
        eng = V._engine()
        from skgstat_b200._exact import sq_threshold_bits
        assert_array_equal(eng.hist_counts(sq_threshold_bits(edges)), cnt)
        Vd = skg.Variogram(c, v, estimator="dowd", n_lags=n_lags, maxlag=maxlag, fit_method=None)
        assert_array_equal(Vd.bin_count, cnt)


def test_nan_observations_and_dowd_at_window_sizes(skg):
    rng = np.random.default_rng(77)
    n = 9000
    c = rng.random((n, 2)) * 1000.0
    v = rng.normal(size=n)
    edges = np.linspace(0, 600.0, 4)[1:]
    # exact per-lag medians from materialised classes (row blocks)
    c2 = ov.coords_2d(c)
    classes = [[] for _ in edges]
    for i0 in range(0, n - 1, 256):
        d, dz, _, _ = ov._row_block(c2, v, i0, min(n - 1, i0 + 256))
        g = ov.lag_groups(d, edges)
        for k in range(len(edges)):
            classes[k].append(dz[g == k])
    med = np.array([np.median(np.concatenate(k)) for k in classes])
    Vd = skg.Variogram(c, v, estimator="dowd", n_lags=3, maxlag=600.0, fit_method=None)
    assert_allclose(Vd.experimental, 2.198 * med ** 2 / 2.0, rtol=1e-12)
