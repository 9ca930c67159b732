"""GPU: BASELINE.json's full-size configurations checked through size-independent
properties (the CPU oracle cannot hold N=100k..200k pairs vectors):

  config 3  N=200 000, cressie and dowd          (2e10 pairs)
  config 4  N=100 000, DirectionalVariogram az=45 tol=22.5, uniform bins
  config 5  OrdinaryKriging on a 2000x2000 grid from N=10 000 samples, max 64, exponential

Properties: pair-count conservation, exact quantile structure of median / uniform bins,
scaling laws of the estimators, permutation invariance (exact counts), directional
symmetry, kriging exactness at the samples / constant-field reproduction / translation
invariance, plus spot checks of sub-blocks against the row-blocked oracle.
"""
import os
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import kriging as okr
from oracle import variogram as ov
from oracle.fields import gaussian_field, random_points

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def skg():
    import skgstat_b200
    skgstat_b200._lib.load()
    return skgstat_b200


@pytest.fixture(scope="module")
def field200k():
    c = random_points(200000, 10000.0, 2, seed=0)
    return c, gaussian_field(c, 1000.0, seed=0)


def test_config3_cressie_and_dowd_200k(skg, field200k):
    c, v = field200k
    n = len(c)
    P = n * (n - 1) // 2
    Vc = skg.Variogram(c, v, estimator="cressie", n_lags=25, maxlag="median", fit_method=None)
    cnt = Vc.bin_count
    # median maxlag: the classes hold every pair strictly below the median distance
    assert abs(int(cnt.sum()) - P // 2) <= 1
    assert np.all(cnt > 0) and np.all(np.diff(cnt[:12]) > 0)  # 2-D: counts grow with short lags
    assert np.all(np.isfinite(Vc.experimental)) and np.all(Vc.experimental > 0)
    # same classes, other estimators: identical counts (three different kernels)
    Vm = skg.Variogram(c, v, estimator="matheron", n_lags=25, maxlag="median", fit_method=None)
    Vd = skg.Variogram(c, v, estimator="dowd", n_lags=25, maxlag="median", fit_method=None)
    assert_array_equal(Vm.bins, Vc.bins)
    assert_array_equal(Vm.bin_count, cnt)
    assert_array_equal(Vd.bin_count, cnt)
    # scaling laws: matheron and dowd ~ a^2, cressie ~ a^2 too ((mean sqrt|a dz|)^4)
    a = 3.0
    Vs = skg.Variogram(c, a * v - 2.0, estimator="cressie", n_lags=25, maxlag="median",
                       fit_method=None)
    assert_array_equal(Vs.bin_count, cnt)
    assert_allclose(Vs.experimental, a * a * Vc.experimental, rtol=1e-11)
    Vds = skg.Variogram(c, a * v, estimator="dowd", n_lags=25, maxlag="median", fit_method=None)
    assert_allclose(Vds.experimental, a * a * Vd.experimental, rtol=1e-12)
    # a Gaussian field: the robust estimators agree with matheron to within 20 % (noise component differs)
    assert_allclose(Vc.experimental, Vm.experimental, rtol=0.2)
    assert_allclose(Vd.experimental, Vm.experimental, rtol=0.2)
    # permutation invariance: counts exact, sums to rounding
    perm = np.random.default_rng(1).permutation(n)
    Vp = skg.Variogram(c[perm], v[perm], estimator="matheron", n_lags=25, maxlag="median",
                       fit_method=None)
    assert Vp.maxlag == Vm.maxlag
    assert_array_equal(Vp.bin_count, cnt)
    assert_allclose(Vp.experimental, Vm.experimental, rtol=1e-12)


def test_config3_subsample_against_oracle(skg, field200k):
    """The first 3000 points of the 200k field: every kernel against the dense oracle."""
    c, v = field200k
    c, v = c[:3000], v[:3000]
    for est in ("cressie", "dowd"):
        ref = ov.dense_variogram(c, v, est, "even", 25, "median")
        V = skg.Variogram(c, v, estimator=est, n_lags=25, maxlag="median", fit_method=None)
        assert_array_equal(V.bins, ref["bins"])
        assert_array_equal(V.bin_count, ref["bin_count"])
        assert_allclose(V.experimental, ref["experimental"], rtol=1e-12)


def test_config4_directional_100k(skg):
    n = 100000
    c = random_points(n, 10000.0, 2, seed=4)
    v = gaussian_field(c, 800.0, seed=4)
    P = n * (n - 1) // 2
    DV = skg.DirectionalVariogram(c, v, azimuth=45, tolerance=22.5, bin_func="uniform",
                                  n_lags=10, fit_method=None)
    cnt = DV.bin_count
    tot = int(cnt.sum())
    # uniform bins: equal counts up to the percentile interpolation (last edge = max, excluded)
    assert cnt.max() - cnt.min() <= 2
    # the mask (cone of +-11.25 deg, band q33/2) keeps a small fraction of the pairs
    assert 0 < tot < 0.13 * P  # the cone of +-11.25 deg keeps <= 12.5 % of the pairs, the band fewer
    assert np.all(np.diff(DV.bins) > 0)
    # the pair direction is a line, not a ray: azimuth and azimuth-180 select the same pairs
    DV2 = skg.DirectionalVariogram(c, v, azimuth=-135, tolerance=22.5, bin_func="uniform",
                                   n_lags=10, fit_method=None)
    assert DV2.bandwidth == DV.bandwidth
    assert_array_equal(DV2.bins, DV.bins)
    assert_array_equal(DV2.bin_count, cnt)
    assert_allclose(DV2.experimental, DV.experimental, rtol=1e-12)
    # mirrored coordinates with mirrored azimuth: same pairs again
    cm = c.copy()
    cm[:, 1] = -cm[:, 1]
    DV3 = skg.DirectionalVariogram(cm, v, azimuth=-45, tolerance=22.5, bin_func="uniform",
                                   n_lags=10, fit_method=None)
    assert_array_equal(DV3.bin_count, cnt)
    assert_allclose(DV3.bins, DV.bins, rtol=1e-15)
    # sub-sample against the dense oracle (same settings)
    cs, vs = c[:2500], v[:2500]
    ref = ov.dense_directional(cs, vs, "matheron", "uniform", 10, None, 45, 22.5, "q33", "triangle")
    DVs = skg.DirectionalVariogram(cs, vs, azimuth=45, tolerance=22.5, bin_func="uniform",
                                   n_lags=10, fit_method=None)
    assert_array_equal(DVs.bins, ref["bins"])
    assert_array_equal(DVs.bin_count, ref["bin_count"])
    assert_allclose(DVs.experimental, ref["experimental"], rtol=1e-12)


def test_config5_kriging_grid_2000x2000(skg):
    c = random_points(10000, 1000.0, 2, seed=5)
    v = gaussian_field(c, 100.0, seed=5)
    descr = dict(model="exponential", effective_range=300.0, sill=4.2, nugget=0.0,
                 dist_func="euclidean")
    ok = skg.OrdinaryKriging(descr, min_points=5, max_points=64, coordinates=c, values=v)
    gx, gy = np.mgrid[0:1000:2000j, 0:1000:2000j].reshape(2, -1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        z = ok.transform(gx, gy)
    s2 = ok.sigma
    assert z.shape == (4000000,) and not np.isnan(z).any() and not np.isnan(s2).any()
    assert ok.no_points_error == 0 and ok.singular_error == 0
    # kriging is a weighted mean with weights summing to one, variance within [0, ~sill]
    assert z.min() > v.min() - 3.0 and z.max() < v.max() + 3.0
    assert s2.min() > -1e-9 and s2.max() < 2 * 4.2
    # spot check 300 random grid nodes against the oracle
    idx = np.random.default_rng(0).choice(z.size, 300, replace=False)
    t = np.column_stack((gx[idx], gy[idx]))
    zr, sr, _ = okr.krige(c, v, t, "exponential", 300.0, 4.2, 0.0, None, 5, 64)
    assert_allclose(z[idx], zr, rtol=1e-9)
    assert_allclose(s2[idx], sr, rtol=1e-9, atol=1e-9 * 4.2)
    # exact interpolator (no nugget): at the samples z = value and sigma = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        zs = ok.transform(c[:2000, 0], c[:2000, 1])
    assert_allclose(zs, v[:2000], rtol=1e-9)
    assert np.abs(ok.sigma).max() < 1e-8
    # constant field is reproduced (weights sum to one); translation invariance
    okc = skg.OrdinaryKriging(descr, min_points=5, max_points=64, coordinates=c,
                              values=np.full(len(c), 7.25))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        zc = okc.transform(gx[::997], gy[::997])
    assert_allclose(zc, 7.25, rtol=1e-12)
    okt = skg.OrdinaryKriging(descr, min_points=5, max_points=64, coordinates=c + 12345.0,
                              values=v)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        zt = okt.transform(gx[::997] + 12345.0, gy[::997] + 12345.0)
    assert_allclose(zt, z[::997], rtol=1e-7)


# ---- full-size pins: oracle-only golden vectors (oracle/make_fullsize_golden.py) ----------------

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize.npz")


@pytest.fixture(scope="module")
def fullsize_golden():
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/fullsize.npz not generated")
    return np.load(GOLDEN)


def test_config3_200k_pinned_to_cpu_oracle(skg, field200k, fullsize_golden):
    """N = 200 000 (2e10 pairs): the exact median maxlag, bit-exact edges and per-lag pair counts and
    1e-12 semivariances of matheron / cressie / dowd against the C oracle's sweep over ALL pairs
    (edges from the oracle's own exact select, integer counts, long-double sums)."""
    g = fullsize_golden
    c, v = field200k
    assert int(g["c3_n"]) == len(c)
    for est in ("matheron", "cressie", "dowd"):
        V = skg.Variogram(c, v, estimator=est, n_lags=25, maxlag="median", fit_method=None)
        assert V.maxlag == float(g["c3_maxlag"])
        assert_array_equal(V.bins, g["c3_bins"])
        assert_array_equal(V.bin_count, g["c3_bin_count"])
        assert_allclose(V.experimental, g["c3_" + est], rtol=1e-12, atol=0)


def test_config4_100k_directional_pinned_to_cpu_oracle(skg, fullsize_golden):
    """N = 100 000 directional (azimuth 45, tolerance 22.5, q33 bandwidth, uniform bins): bandwidth,
    bit-exact uniform edges and counts, 1e-12 semivariances against the C oracle, whose mask is the
    reference's literal arccos / sin formula evaluated for every one of the 5e9 pairs."""
    g = fullsize_golden
    n = int(g["c4_n"])
    c = random_points(n, 10000.0, 2, seed=4)
    v = gaussian_field(c, 800.0, seed=4)
    DV = skg.DirectionalVariogram(c, v, azimuth=45, tolerance=22.5, bin_func="uniform", n_lags=10,
                                  fit_method=None)
    assert DV.bandwidth == float(g["c4_bandwidth"])
    assert_array_equal(DV.bins, g["c4_bins"])
    assert_array_equal(DV.bin_count, g["c4_bin_count"])
    assert_allclose(DV.experimental, g["c4_matheron"], rtol=1e-12, atol=0)
