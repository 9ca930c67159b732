"""GPU parity tests (run on the B200 box): the CUDA path, called through the
reference-shaped classes and the C ABI, against
  * the reference's own outputs (tests/golden/*.npz, see oracle/make_golden.py),
  * the CPU oracle on seeded inputs,
  * size-independent properties at BASELINE.json's full size.
Tolerances follow BASELINE.json: bin edges and counts bit-exact, semivariances
rtol 1e-12, kriging rtol 1e-9.
"""
import pickle
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import variogram as ov
from oracle.fields import gaussian_field, random_points

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def skg():
    import skgstat_b200
    skgstat_b200._lib.load()
    return skgstat_b200


def _check_case(V, g, case):
    assert_array_equal(V.bins, g.out(case, "bins"), err_msg=str(case))
    assert_array_equal(V.bin_count, g.out(case, "bin_count"), err_msg=str(case))
    assert_allclose(V.experimental, g.out(case, "experimental"), rtol=1e-12, equal_nan=True,
                    err_msg=str(case))
    ml = g.out(case, "maxlag")
    if np.isnan(ml):
        assert V.maxlag is None
    elif case["maxlag"] == "mean":
        assert V.maxlag == ml
    else:
        assert V.maxlag == ml, str(case)


def test_variogram_reference_golden(skg, golden_variogram):
    g = golden_variogram
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for case in g.manifest:
            c, v = g.inputs(case["dataset"])
            V = skg.Variogram(c, v, estimator=case["estimator"], bin_func=case["bin_func"],
                              n_lags=case["n_lags"], maxlag=case["maxlag"])
            _check_case(V, g, case)
            # CPU model fit on identical (bins, experimental) lands on the same optimum
            assert_allclose(V.cof, g.out(case, "cof"), rtol=1e-5, atol=1e-7, err_msg=str(case))


def test_reference_known_answers(skg):
    """skgstat/tests/test_variogram.py:677-685, :229-245, :105-124, :481-487."""
    np.random.seed(42)
    c = np.random.gamma(10, 4, (30, 2))
    np.random.seed(42)
    v = np.random.normal(10, 4, 30)
    V = skg.Variogram(c, v)
    assert isinstance(V.bin_count, np.ndarray)
    assert_array_equal(V.bin_count, [22, 54, 87, 65, 77, 47, 46, 24, 10, 2])
    for x, y in zip(V.parameters, [7.122, 13.966, 0]):
        assert abs(x - y) < 1e-3
    old = V.bin_count
    V.bin_func = "uniform"
    assert not np.array_equal(V.bin_count, old)
    old = V.bin_count
    V.maxlag = 25
    assert not np.array_equal(V.bin_count, old)
    V.n_lags = 5
    assert len(V.bin_count) == 5
    V = skg.Variogram(c, v, n_lags=4)
    np.testing.assert_array_almost_equal(V.bins, [10.58, 21.15, 31.73, 42.3], decimal=2)
    V.set_bin_func("uniform")
    np.testing.assert_array_almost_equal(V.bins, [10.25, 16.21, 22.71, 42.3], decimal=2)
    V.bin_func = "even"
    np.testing.assert_array_almost_equal(V.bins, [10.58, 21.15, 31.73, 42.3], decimal=2)
    V = skg.Variogram(c, v, maxlag="mean", n_lags=4)
    np.testing.assert_array_almost_equal(V.bins, [4.23, 8.46, 12.69, 16.91], decimal=2)
    V = skg.Variogram(c, v)
    V.maxlag = 0.6
    assert V.maxlag == np.max(V.distance) * 0.6
    assert abs(V.maxlag - 25.38) < 1e-2


def test_pancake_known_answers(skg, golden_variogram):
    """BASELINE.md section 3."""
    c, v = golden_variogram.inputs("pancake300")
    V = skg.Variogram(c, v)
    assert V.bins[0] == float.fromhex("0x1.03093b2e17b39p+6")
    assert V.bins[9] == float.fromhex("0x1.43cb89f99da07p+9")
    assert_array_equal(V.bin_count, [2153, 5570, 7720, 8573, 7849, 6165, 4216, 2010, 527, 66])
    assert_allclose(V.experimental[[0, 9]], [480.1110078959591, 1105.5075757575758], rtol=1e-13)
    assert_allclose(V.cof, [353.63831427921605, 1512.2358386536657], rtol=1e-6)
    V2 = skg.Variogram(c, v, model="exponential", n_lags=25, maxlag="median")
    assert V2.maxlag == 246.31382346127907
    assert int(V2.bin_count.sum()) == 22425
    assert_array_equal(V2.bin_count[:6], [57, 176, 248, 378, 430, 519])


@pytest.mark.parametrize("n", [2, 3, 255, 256, 257, 513, 1500])
@pytest.mark.parametrize("estimator", ["matheron", "cressie", "dowd"])
def test_against_oracle_ragged_sizes(skg, n, estimator):
    c = random_points(n, 100.0, 2, seed=n)
    v = gaussian_field(c, 10.0, seed=n)
    for bf, nl, ml in (("even", 7, None), ("uniform", 5, "median"), ("even", 25, 0.5)):
        ref = ov.dense_variogram(c, v, estimator, bf, nl, ml)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            V = skg.Variogram(c, v, estimator=estimator, bin_func=bf, n_lags=nl, maxlag=ml,
                              fit_method=None)
        assert_array_equal(V.bins, ref["bins"])
        assert_array_equal(V.bin_count, ref["bin_count"])
        assert_allclose(V.experimental, ref["experimental"], rtol=1e-12, equal_nan=True)


def test_many_lags_and_3d_and_1d(skg):
    c = random_points(900, 50.0, 3, seed=9)
    v = gaussian_field(c, 9.0, seed=9)
    for nl in (1, 64, 100, 200):  # 100/200 exceed the private-accumulator budget -> atomic path
        for est in ("matheron", "cressie"):
            ref = ov.dense_variogram(c, v, est, "even", nl, None)
            V = skg.Variogram(c, v, estimator=est, n_lags=nl, fit_method=None)
            assert_array_equal(V.bins, ref["bins"])
            assert_array_equal(V.bin_count, ref["bin_count"])
            assert_allclose(V.experimental, ref["experimental"], rtol=1e-12, equal_nan=True)
    c1 = random_points(400, 80.0, 1, seed=4)[:, 0]
    v1 = gaussian_field(c1[:, None], 6.0, seed=4)
    ref = ov.dense_variogram(c1, v1, "matheron", "uniform", 6, "median")
    V = skg.Variogram(c1, v1, bin_func="uniform", n_lags=6, maxlag="median", fit_method=None)
    assert V.dim == 1
    assert_array_equal(V.bins, ref["bins"])
    assert_array_equal(V.bin_count, ref["bin_count"])


def test_ties_duplicates_and_custom_edges(skg):
    gx, gy = np.meshgrid(np.arange(40), np.arange(40))
    c = np.column_stack((gx.ravel(), gy.ravel())).astype(float)
    c = np.vstack((c, c[:25]))  # exact duplicates -> zero distances
    v = gaussian_field(c, 5.0, seed=1)
    for bf in ("even", "uniform"):
        for est in ("matheron", "dowd"):
            ref = ov.dense_variogram(c, v, est, bf, 12, None)
            V = skg.Variogram(c, v, estimator=est, bin_func=bf, n_lags=12, fit_method=None)
            assert_array_equal(V.bins, ref["bins"])
            assert_array_equal(V.bin_count, ref["bin_count"])
            assert_allclose(V.experimental, ref["experimental"], rtol=1e-12, equal_nan=True)
    edges = [1.0, 1.0, 2.5, 5.0, 5.0, 5.0, 17.0, 30.0]  # repeated edges: empty classes
    d = ov.distance(c)
    g = ov.lag_groups(d, np.asarray(edges))
    V = skg.Variogram(c, v, bin_func=edges, fit_method=None)
    assert_array_equal(V.bin_count, [(g == k).sum() for k in range(len(edges))])
    assert V.n_lags == len(edges) and V.maxlag == 30.0


def test_materialising_accessors(skg, golden_variogram):
    c, v = golden_variogram.inputs("rnd400")
    V = skg.Variogram(c, v, n_lags=9, maxlag="median", fit_method=None)
    assert_array_equal(V.distance, ov.distance(c))
    assert_array_equal(V.pairwise_diffs, ov.pairwise_diffs(v))
    assert_array_equal(V.lag_groups(), ov.lag_groups(ov.distance(c), V.bins))
    sizes = [k.size for k in V.lag_classes()]
    assert_array_equal(sizes, V.bin_count)
    assert V.distance_matrix.shape == (400, 400)


def test_setters_invalidate_and_pickle(skg, golden_variogram):
    c, v = golden_variogram.inputs("rnd400")
    V = skg.Variogram(c, v, n_lags=8)
    e0 = V.experimental
    V.estimator = "cressie"
    assert not np.allclose(V.experimental, e0)
    V.set_values(v * 2.0)
    V.estimator = "matheron"
    assert_allclose(V.experimental, 4.0 * e0, rtol=1e-12)
    V2 = pickle.loads(pickle.dumps(V))
    assert_array_equal(V2.bins, V.bins)
    assert_allclose(V2.experimental, V.experimental, rtol=0, atol=0)
    V3 = V.clone()
    V3.n_lags = 4
    assert len(V3.bins) == 4 and len(V.bins) == 8
    with pytest.raises(ValueError):
        V.n_lags = -3
    with pytest.raises(ValueError):
        V.set_estimator("notanestimator")


def test_full_size_properties(skg):
    """BASELINE.json config 2 (N=20k, matheron, 25 lags, maxlag=median): exact counts
    against the row-blocked oracle, linearity, and run-to-run reproducibility."""
    n = 20000
    c = random_points(n, 1000.0, 2, seed=0)
    v = gaussian_field(c, 100.0, seed=0)
    V = skg.Variogram(c, v, n_lags=25, maxlag="median", fit_method=None)
    cnt = V.bin_count
    P = n * (n - 1) // 2
    # median maxlag: exactly half of the pairs (rounded down) lie strictly below it ...
    assert abs(int(cnt.sum()) - P // 2) <= 1
    blk = ov.blocked_hist(c, v, V.bins, block=200)
    assert_array_equal(cnt, blk["count"])
    assert_allclose(V.experimental, ov.finalize(blk["count"], blk["sum_sq"], blk["sum_sqrt"],
                                                "matheron"), rtol=1e-12)
    e1 = V.experimental
    V.preprocessing(force=True)
    assert_array_equal(V.experimental, e1)  # fixed-order reduction: bitwise reproducible
    Vs = skg.Variogram(c, 3.0 * v + 7.0, n_lags=25, maxlag="median", fit_method=None)
    assert_array_equal(Vs.bin_count, cnt)
    assert_allclose(Vs.experimental, 9.0 * e1, rtol=1e-11)


def test_windowed_select_paths_are_exact(skg, monkeypatch):
    """The large-problem paths (fine histogram + gather inside a crowded distance cell instead of
    the direct gather; sample-derived windows for Dowd's per-lag medians) forced on at a size the
    dense oracle can check."""
    from skgstat_b200.engine import PairEngine
    monkeypatch.setattr(PairEngine, "DIRECT_GATHER_LIMIT", 0)
    monkeypatch.setattr(PairEngine, "WINDOWED_MEDIAN_MIN_PAIRS", 0)
    c = random_points(9000, 1000.0, 2, seed=17)
    v = gaussian_field(c, 80.0, seed=17)
    for est, bf, nl, ml in (("dowd", "even", 20, "median"), ("matheron", "uniform", 8, None),
                            ("dowd", "uniform", 6, 0.4)):
        ref = ov.dense_variogram(c, v, est, bf, nl, ml)
        V = skg.Variogram(c, v, estimator=est, bin_func=bf, n_lags=nl, maxlag=ml, fit_method=None)
        assert_array_equal(V.bins, ref["bins"])
        assert_array_equal(V.bin_count, ref["bin_count"])
        assert_allclose(V.experimental, ref["experimental"], rtol=1e-12, equal_nan=True)
    # heavy ties: lattice coordinates and quantised values
    gx, gy = np.meshgrid(np.arange(70.0), np.arange(70.0))
    cl = np.column_stack((gx.ravel(), gy.ravel()))
    vl = np.round(gaussian_field(cl, 9.0, seed=3), 1)
    ref = ov.dense_variogram(cl, vl, "dowd", "even", 12, "median")
    V = skg.Variogram(cl, vl, estimator="dowd", n_lags=12, maxlag="median", fit_method=None)
    assert_array_equal(V.bins, ref["bins"])
    assert_array_equal(V.bin_count, ref["bin_count"])
    assert_allclose(V.experimental, ref["experimental"], rtol=1e-12, equal_nan=True)


def test_nan_observations_follow_numpy_nan_rules(skg):
    """NaN values: matheron / cressie propagate NaN like np.sum, dowd uses np.nanmedian
    (estimators.py:178) and bin_count still counts every pair of the lag class."""
    c = random_points(700, 100.0, 2, seed=31)
    v = gaussian_field(c, 10.0, seed=31)
    v[[5, 77, 300]] = np.nan
    for est in ("matheron", "dowd"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = ov.dense_variogram(c, v, est, "even", 10, "median")
            V = skg.Variogram(c, v, estimator=est, n_lags=10, maxlag="median", fit_method=None)
            exp = V.experimental
        assert_array_equal(V.bins, ref["bins"])
        assert_array_equal(V.bin_count, ref["bin_count"])
        assert_allclose(exp, ref["experimental"], rtol=1e-12, equal_nan=True)


def test_windowed_dowd_kernels_stress(skg, monkeypatch):
    """The large-N Dowd path (merged per-lag |dz| windows, the shared-window sweep loop, the
    value-first gather kernel, the pick kernels) forced on at sizes the dense oracle checks:
    ragged N, many narrow lag classes (groups spanning more than three), clustered custom edges,
    a field without spatial structure (every class gets the same window), heavy ties and
    constant values (degenerate windows)."""
    from skgstat_b200.engine import PairEngine
    monkeypatch.setattr(PairEngine, "WINDOWED_MEDIAN_MIN_PAIRS", 0)
    monkeypatch.setattr(PairEngine, "SUBSAMPLE_POINTS", 1024)
    rng = np.random.default_rng(5)
    cases = []
    c = random_points(4099, 1000.0, 2, seed=21)
    cases.append((c, gaussian_field(c, 120.0, seed=21), dict(n_lags=25, maxlag="median")))
    cases.append((c, gaussian_field(c, 120.0, seed=22), dict(n_lags=120, maxlag="median")))
    cases.append((c, rng.normal(size=c.shape[0]), dict(n_lags=40, maxlag=0.6)))           # pure nugget
    cases.append((c, gaussian_field(c, 60.0, seed=23),
                  dict(bin_func=[5.0, 5.5, 6.0, 50.0, 50.000001, 200.0, 400.0, 401.0, 700.0], maxlag=None)))
    c2 = random_points(3001, 500.0, 2, seed=24)
    cases.append((c2, np.round(gaussian_field(c2, 70.0, seed=24), 1), dict(n_lags=15, maxlag="median")))  # ties
    cases.append((c2, np.full(c2.shape[0], 3.25), dict(n_lags=10, maxlag="median")))                      # constant
    gx, gy = np.meshgrid(np.arange(60.0), np.arange(55.0))
    cl = np.column_stack((gx.ravel(), gy.ravel()))
    cases.append((cl, np.round(gaussian_field(cl, 9.0, seed=3), 1), dict(n_lags=12, maxlag="median")))    # lattice
    for coords, values, kw in cases:
        if isinstance(kw.get("bin_func"), list):   # custom edges (Variogram.py:1119-1131): classes straight from them
            edges = np.asarray(kw["bin_func"], dtype=float)
            d, diffs = ov.distance(coords), ov.pairwise_diffs(values)
            g = ov.lag_groups(d, edges)
            cls = [diffs[np.where(g == i)] for i in range(len(edges))]
            ref = dict(bins=edges, bin_count=np.array([x.size for x in cls]),
                       experimental=np.array([ov.dowd(x) for x in cls]))
        else:
            ref = ov.dense_variogram(coords, values, "dowd", "even", kw.get("n_lags", 10), kw.get("maxlag"))
        V = skg.Variogram(coords, values, estimator="dowd", fit_method=None, **kw)
        assert_array_equal(V.bins, ref["bins"])
        assert_array_equal(V.bin_count, ref["bin_count"])
        assert_allclose(V.experimental, ref["experimental"], rtol=1e-12, atol=0.0, equal_nan=True)
