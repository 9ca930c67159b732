"""GPU parity for SURVEY.md section 8(f) row N1: cross-variograms, multivariate='aggregate', the
cross_variograms grid, the batched pair sweep (many value vectors over shared coordinates) and
the Monte-Carlo propagate utility - against the reference's outputs (tests/golden/multivariate.npz,
oracle/make_golden.py::make_multivariate) and the CPU oracle.  Tolerances as everywhere: bins and
counts bit-exact, semivariances rtol 1e-12."""
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


def _cols(vals, mode):
    return vals[:, :2] if mode == "cross" else vals


def test_cross_and_aggregate_match_reference(skg, golden_mv):
    g = golden_mv
    n = 0
    for case in g.manifest:
        if case["kind"] != "variogram":
            continue
        c, vals = g.inputs(case["dataset"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            V = skg.Variogram(c, _cols(vals, case["mode"]), estimator=case["estimator"],
                              bin_func=case["bin_func"], n_lags=case["n_lags"], maxlag=case["maxlag"],
                              multivariate=case["mode"], fit_method=None)
        assert_array_equal(V.bins, g.out(case, "bins"), err_msg=str(case))
        assert_array_equal(V.bin_count, g.out(case, "bin_count"), err_msg=str(case))
        assert_allclose(V.experimental, g.out(case, "experimental"), rtol=1e-12, equal_nan=True,
                        err_msg=str(case))
        assert V.is_cross_variogram == (case["mode"] == "cross")
        assert V.is_agg_variogram == (case["mode"] == "aggregate")
        n += 1
    assert n == 54


def test_pairwise_diffs_of_multicolumn_values(skg, golden_mv):
    """Variogram.pairwise_diffs (materialised on the device) carries the combined difference."""
    g = golden_mv
    seen = set()
    for case in g.manifest:
        if case["kind"] != "variogram" or (case["dataset"], case["mode"]) in seen:
            continue
        seen.add((case["dataset"], case["mode"]))
        c, vals = g.inputs(case["dataset"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            V = skg.Variogram(c, _cols(vals, case["mode"]), multivariate=case["mode"], fit_method=None)
        d = V.pairwise_diffs
        assert d.ndim == 1 and d.size == len(c) * (len(c) - 1) // 2
        assert_array_equal(d, ov.pairwise_diffs_multi(_cols(vals, case["mode"]), case["mode"]))
        chk = g.out(case, "diff_checksum")
        assert_allclose(d.sum(), chk[0], rtol=1e-13)
        assert d.max() == chk[1]
    assert len(seen) == 6


def test_directional_cross_variograms_match_reference(skg, golden_mv):
    g = golden_mv
    n = 0
    for case in g.manifest:
        if case["kind"] != "directional":
            continue
        c, vals = g.inputs(case["dataset"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            V = skg.DirectionalVariogram(c, vals[:, :2], estimator=case["estimator"], n_lags=case["n_lags"],
                                         maxlag=case["maxlag"], azimuth=case["azimuth"],
                                         tolerance=case["tolerance"],
                                         directional_model=case["directional_model"], fit_method=None)
        assert_array_equal(V.bins, g.out(case, "bins"))
        assert_array_equal(V.bin_count, g.out(case, "bin_count"))
        assert_allclose(V.experimental, g.out(case, "experimental"), rtol=1e-12, equal_nan=True)
        n += 1
    assert n == 4


def test_cross_variograms_grid_matches_reference(skg, golden_mv):
    from skgstat_b200.util.cross_variogram import cross_variograms
    g = golden_mv
    case = [m for m in g.manifest if m["kind"] == "grid"][0]
    c, vals = g.inputs(case["dataset"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mat = cross_variograms(c, vals, n_lags=case["n_lags"], maxlag=case["maxlag"], fit_method=None)
    exp = g.out(case, "experimental")
    cnt = g.out(case, "bin_count")
    is_cross = g.out(case, "is_cross")
    assert len(mat) == 3 and all(len(r) == 3 for r in mat)
    for i in range(3):
        for j in range(3):
            assert_array_equal(mat[i][j].bin_count, cnt[i, j])
            assert_allclose(mat[i][j].experimental, exp[i, j], rtol=1e-12)
            assert bool(mat[i][j].is_cross_variogram) == bool(is_cross[i, j])
    # one MetricSpace (one device copy of the coordinates) behind all nine instances
    assert len({id(mat[i][j]._X) for i in range(3) for j in range(3)}) == 1
    # directional kwargs select the directional base class (util/cross_variogram.py:64-67)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dmat = cross_variograms(c[:120], vals[:120, :2], azimuth=90, fit_method=None)
    assert isinstance(dmat[0][1], skg.DirectionalVariogram)


def test_reference_cross_variogram_unit_tests(skg):
    """tests/test_variogram.py:1589-1665 restated against the drop-in class."""
    np.random.seed(42)
    c = np.random.gamma(10, 4, (100, 2))
    np.random.seed(42)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        v = np.random.multivariate_normal([10.5, 5.6], [[1.5, 3.2], [3.2, 1.5]], size=100)
    np.random.seed(42)
    with pytest.raises(ValueError, match="values has to be 1d"):
        skg.Variogram(c, np.random.normal(10, 4, (100, 3, 3)))
    with pytest.raises(ValueError, match="create a grid of cross-variograms"):
        skg.Variogram(c, np.random.normal(10, 3, (100, 3)))
    vario = skg.Variogram(c, v, maxlag="median")
    assert vario.is_cross_variogram
    assert vario.pairwise_diffs.ndim == 1
    assert_allclose(v[:, 1], vario._co_variable)
    assert_array_equal(vario.values, v[:, 0])
    dv = skg.DirectionalVariogram(c, v, maxlag="median")
    assert dv.is_cross_variogram and dv.pairwise_diffs.ndim == 1
    assert_allclose(v[:, 1], dv._co_variable)
    with pytest.raises(NotImplementedError):
        skg.Variogram(c, v, multivariate="nope")


@pytest.mark.parametrize("n,dim,k,est", [(1500, 2, 1, "matheron"), (1500, 2, 5, "matheron"),
                                          (900, 3, 4, "cressie"), (2100, 2, 9, "cressie"),
                                          (257, 2, 3, "matheron")])
def test_batched_sweep_equals_one_sweep_per_vector(skg, n, dim, k, est):
    """skg_pair_hist_batch against the oracle and against the single-vector kernel."""
    c = random_points(n, 100.0, dim, seed=n)
    base = gaussian_field(c, 15.0, seed=n)
    rng = np.random.default_rng(n + k)
    draws = np.vstack([base + rng.normal(0, 0.3, n) for _ in range(k)])
    V = skg.Variogram(c, base, estimator=est, n_lags=14, maxlag="median", fit_method=None)
    got = V.experimental_batch(draws)
    assert got.shape == (k, 14)
    for q in range(k):
        ref = ov.dense_variogram(c, draws[q], est, "even", 14, "median")
        assert_array_equal(V.bins, ref["bins"])
        assert_allclose(got[q], ref["experimental"], rtol=1e-12, equal_nan=True)
        single = skg.Variogram(V._X, draws[q], estimator=est, n_lags=14, maxlag="median", fit_method=None)
        assert_array_equal(single.bin_count, V.bin_count)
        assert_allclose(got[q], single.experimental, rtol=1e-13)
    # the batch leaves the instance itself untouched
    assert_array_equal(V.values, base)
    assert_allclose(V.experimental, ov.dense_variogram(c, base, est, "even", 14, "median")["experimental"],
                    rtol=1e-12)


def test_batched_sweep_directional_and_fallbacks(skg):
    c = random_points(1200, 100.0, 2, seed=3)
    base = gaussian_field(c, 12.0, seed=3)
    rng = np.random.default_rng(0)
    draws = np.vstack([base + rng.normal(0, 0.5, len(base)) for _ in range(6)])
    kw = dict(azimuth=30, tolerance=40.0, bandwidth=25.0, n_lags=9, maxlag="median", fit_method=None)
    DV = skg.DirectionalVariogram(c, base, **kw)
    got = DV.experimental_batch(draws)
    for q in range(6):
        one = skg.DirectionalVariogram(c, draws[q], **kw)
        assert_allclose(got[q], one.experimental, rtol=1e-13)
    # dowd has no batched kernel: one pass per vector, same answer as separate instances
    D = skg.Variogram(c, base, estimator="dowd", n_lags=9, maxlag="median", fit_method=None)
    gd = D.experimental_batch(draws[:3])
    for q in range(3):
        one = skg.Variogram(c, draws[q], estimator="dowd", n_lags=9, maxlag="median", fit_method=None)
        assert_allclose(gd[q], one.experimental, rtol=1e-13)
    assert_array_equal(D.values, base)
    # more lags than the batched records hold: per-vector path
    many = skg.Variogram(c, base, n_lags=120, fit_method=None)
    gm = many.experimental_batch(draws[:2])
    one = skg.Variogram(c, draws[1], n_lags=120, fit_method=None)
    assert_allclose(gm[1], one.experimental, rtol=1e-13, equal_nan=True)
    with pytest.raises(ValueError):
        D.experimental_batch(draws[:, :10])


def test_propagate_matches_reference(skg, golden_mv):
    """util/uncertainty.py: same random stream, same confidence intervals."""
    from skgstat_b200.util.uncertainty import propagate
    g = golden_mv
    n = 0
    for case in g.manifest:
        if case["kind"] not in ("propagate", "propagate3"):
            continue
        c, vals = g.inputs(case["dataset"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            V = skg.Variogram(c, vals[:, 0], n_lags=case["n_lags"], maxlag=case["maxlag"],
                              model="exponential")
            if case["kind"] == "propagate":
                V.estimator = case["estimator"]
                ci = propagate(V, "values", sigma=case["sigma"], evalf="experimental",
                               num_iter=case["num_iter"], seed=case["seed"])
                assert_allclose(ci, g.out(case, "conf"), rtol=1e-11)
            else:
                res = propagate(V, "values", sigma=case["sigma"],
                                evalf=["experimental", "parameter", "model"], num_iter=case["num_iter"],
                                seed=case["seed"], q=case["q"], eval_at=case["eval_at"])
                assert len(res) == 3
                assert_allclose(res[0], g.out(case, "conf_exp"), rtol=1e-11)
                assert_allclose(res[1], g.out(case, "conf_par"), rtol=1e-5, atol=1e-8)
                assert_allclose(res[2], g.out(case, "conf_model"), rtol=1e-5)
        n += 1
    assert n == 3


def test_propagate_generic_path_and_bounds(skg):
    from skgstat_b200.util.uncertainty import propagate
    c = random_points(300, 50.0, 2, seed=5)
    v = gaussian_field(c, 8.0, seed=5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        V = skg.Variogram(c, v, n_lags=8, maxlag="median")
        ci = propagate(V, "values", sigma=0.1, num_iter=10, seed=3, use_bounds=True)
        assert ci.shape == (8, 3)
        assert np.all(ci[:, 0] <= ci[:, 1]) and np.all(ci[:, 1] <= ci[:, 2])
        # cressie draws through the same batched sweep
        V.estimator = "cressie"
        ci2 = propagate(V, "values", sigma=0.1, num_iter=6, seed=3, q=50)
        assert ci2.shape == (8, 3) and np.all(np.isfinite(ci2))
        # any other source runs the reference's loop, which (as in the reference, whose params
        # carry no values) cannot construct the variogram: "only 'values' is really supported"
        with pytest.raises(TypeError):
            propagate(V, "maxlag", sigma=0.5, num_iter=2, seed=3)
