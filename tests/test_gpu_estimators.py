"""GPU parity for the minmax / percentile / entropy estimators (SURVEY.md 8(f) N4;
estimators.py:257-367, Variogram.py:2044-2092) against reference outputs
(tests/golden/estimators.npz)."""
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


def test_extra_estimators_match_reference(skg, golden_estimators):
    g = golden_estimators
    for case in g.manifest:
        c, v = g.inputs(case["dataset"])
        kw = dict(case["kwargs"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if case["kind"] == "variogram":
                V = skg.Variogram(c, v, estimator=case["estimator"], bin_func=case["bin_func"],
                                  n_lags=case["n_lags"], maxlag=case["maxlag"], fit_method=None, **kw)
            else:
                V = skg.DirectionalVariogram(c, v, estimator=case["estimator"], n_lags=case["n_lags"],
                                             maxlag=case["maxlag"], azimuth=case["azimuth"],
                                             tolerance=case["tolerance"], bandwidth=case["bandwidth"],
                                             fit_method=None, **kw)
            exp = V.experimental
        assert_array_equal(V.bins, g.out(case, "bins"), err_msg=str(case))
        assert_array_equal(V.bin_count, g.out(case, "bin_count"), err_msg=str(case))
        assert_allclose(exp, g.out(case, "experimental"), rtol=1e-12, equal_nan=True, err_msg=str(case))
    assert len(g.manifest) == 85


def test_value_histogram_and_order_stats_against_materialised_classes(skg):
    from oracle.fields import gaussian_field, random_points
    c = random_points(1100, 100.0, 2, seed=8)
    v = np.round(gaussian_field(c, 10.0, seed=8), 2)     # rounded values: many tied |dz|
    V = skg.Variogram(c, v, n_lags=7, maxlag="median", fit_method=None)
    eng = V._engine()
    edges = V.bins
    d = ov.distance(c)
    dz = ov.pairwise_diffs(v)
    grp = ov.lag_groups(d, edges)
    ve = np.linspace(0.0, dz.max(), 23)
    H = eng.dz_value_histogram(edges, ve)
    cnt, tot = eng.dz_abs_sums(edges)
    vals, totals = eng.dz_order_stats(edges, lambda g, t: [0, t // 3, t - 1] if t else [])
    for k in range(len(edges)):
        x = dz[grp == k]
        assert_array_equal(H[k], np.histogram(x, bins=ve)[0])
        assert cnt[k] == x.size == totals[k]
        assert_allclose(tot[k], x.sum(), rtol=1e-12)
        srt = np.sort(x)
        assert [vals[k][r] for r in (0, x.size // 3, x.size - 1)] == [srt[0], srt[x.size // 3], srt[-1]]


def test_empty_lag_classes_give_nan_not_an_exception(skg):
    """The reference's minmax / percentile raise on an empty lag class (np.nanmax / np.percentile of
    an empty array); the drop-in returns NaN for that class like its other estimators."""
    c = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [10.0, 10.0], [11.0, 10.0]])
    v = np.array([1.0, 2.0, 4.0, 3.0, 5.0])
    for est in ("minmax", "percentile", "entropy"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            V = skg.Variogram(c, v, estimator=est, n_lags=6, fit_method=None)
            exp = V.experimental
        assert (V.bin_count == 0).any()
        assert np.isnan(exp[V.bin_count == 0]).all() or est == "entropy"
        assert np.isfinite(exp[V.bin_count > 0]).all()
    with pytest.raises(NotImplementedError):
        skg.Variogram(c, v, estimator="genton", fit_method=None)
