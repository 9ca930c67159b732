"""GPU parity for the histogram-rule binners (binning.py:82-122) against
reference outputs (tests/golden/binners.npz): edges and counts bit-exact, semivariances 1e-12."""
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def skg():
    import skgstat_b200
    skgstat_b200._lib.load()
    return skgstat_b200


def test_auto_binners_match_reference(skg, golden_binners):
    g = golden_binners
    seen = 0
    for case in g.manifest:
        c, v = g.inputs(case["dataset"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if case["kind"] == "variogram":
                V = skg.Variogram(c, v, bin_func=case["bin_func"], maxlag=case["maxlag"], n_lags=10,
                                  fit_method=None)
            else:
                V = skg.DirectionalVariogram(c, v, bin_func=case["bin_func"], maxlag=case["maxlag"],
                                             azimuth=case["azimuth"], tolerance=case["tolerance"],
                                             bandwidth=case["bandwidth"], fit_method=None)
        assert_array_equal(V.bins, g.out(case, "bins"), err_msg=str(case))
        assert V.n_lags == case["n_lags"]
        assert_array_equal(V.bin_count, g.out(case, "bin_count"), err_msg=str(case))
        assert_allclose(V.experimental, g.out(case, "experimental"), rtol=1e-12, equal_nan=True,
                        err_msg=str(case))
        seen += 1
    assert seen == len(g.manifest) == 128


def test_distance_moments_against_materialised_vector(skg):
    from oracle.fields import gaussian_field, random_points
    from oracle import variogram as ov
    c = random_points(1300, 100.0, 2, seed=2)
    v = gaussian_field(c, 10.0, seed=2)
    V = skg.Variogram(c, v, fit_method=None)
    d = ov.distance(c)
    lim = float(np.median(d))
    n, s1, s2, s3 = V._X.engine().dist_moments(lim)
    dd = d[d <= lim]
    assert n == dd.size
    assert_allclose([s1, s2, s3], [dd.sum(), (dd ** 2).sum(), (dd ** 3).sum()], rtol=1e-12)


def test_auto_binner_limits_and_setters(skg):
    from oracle.fields import gaussian_field, random_points
    c = random_points(3000, 100.0, 2, seed=1)
    v = gaussian_field(c, 10.0, seed=1)
    V = skg.Variogram(c, v, bin_func="sturges", fit_method=None)
    n_sturges = V.n_lags
    assert n_sturges == len(V.bins) == int(np.ceil(np.log2(3000 * 2999 // 2) + 1))
    V.bin_func = "doane"
    assert len(V.bins) >= n_sturges and V.n_lags == len(V.bins)
    # 2121 lag classes: more than one sweep's lag table holds -> consecutive groups of classes
    from oracle import variogram as ov
    big = skg.Variogram(c, v, bin_func="sqrt", fit_method=None)
    d = ov.distance(c)
    ref_edges = np.histogram_bin_edges(d, bins="sqrt")[1:]
    assert_array_equal(big.bins, ref_edges)
    assert big.n_lags == len(ref_edges) == 2121
    g = ov.lag_groups(d, ref_edges)
    assert_array_equal(big.bin_count, np.bincount(g[g >= 0], minlength=len(ref_edges)))
    dz = ov.pairwise_diffs(v)
    with np.errstate(invalid="ignore", divide="ignore"):
        ref_exp = np.bincount(g[g >= 0], weights=dz[g >= 0] ** 2, minlength=len(ref_edges)) / (2.0 * big.bin_count)
    assert_allclose(big.experimental, ref_exp, rtol=1e-12, equal_nan=True)
    for est in ("cressie", "dowd"):
        many = skg.Variogram(c[:1200], v[:1200], estimator=est, n_lags=600, maxlag="median", fit_method=None)
        ref = ov.dense_variogram(c[:1200], v[:1200], est, "even", 600, "median")
        assert_array_equal(many.bin_count, ref["bin_count"])
        assert_allclose(many.experimental, ref["experimental"], rtol=1e-12, equal_nan=True)
    assert_array_equal(many.lag_groups(), ov.lag_groups(ov.distance(c[:1200]), ref["bins"]))
    with pytest.raises(NotImplementedError):
        skg.Variogram(c, v, bin_func="kmeans", fit_method=None)
    with pytest.raises(ValueError):
        skg.Variogram(c, v, bin_func="nonsense", fit_method=None)
