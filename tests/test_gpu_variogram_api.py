"""GPU: the reference's own Variogram unit tests, restated against the drop-in class
(skgstat/tests/test_variogram.py; line numbers in each test).  Same inputs, same
expected numbers, same exception types and messages."""
import os
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_almost_equal
from scipy.spatial.distance import pdist

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def skg():
    import skgstat_b200
    skgstat_b200._lib.load()
    return skgstat_b200


@pytest.fixture()
def cv30():
    np.random.seed(42)
    c = np.random.gamma(10, 4, (30, 2))
    np.random.seed(42)
    v = np.random.normal(10, 4, 30)
    return c, v


@pytest.fixture()
def fitV(skg):
    """TestVariogramFittingProcedure.setUp (:704-714)."""
    np.random.seed(1337)
    c = np.random.gamma(10, 8, (50, 3))
    np.random.seed(1337)
    v = np.random.normal(10, 4, 50)
    return skg.Variogram(c, v, n_lags=5, normalize=False, use_nugget=True), c, v


# ---- TestVariogramInstantiation (:72-218) -----------------------------------------

def test_sparse_standard_settings(skg, cv30):
    V = skg.Variogram(*cv30, maxlag=10000)
    for x, y in zip(V.parameters, [7.122, 13.966, 0]):
        assert abs(x - y) < 1e-3


def test_input_dimensionality(skg):
    c1d = np.random.normal(0, 1, 100)
    c3d = np.random.normal(0, 1, size=(100, 3))
    v = np.random.normal(10, 4, 100)
    assert skg.Variogram(c1d, v).dim == 1
    assert skg.Variogram(c3d, v).dim == 3


def test_pass_median_maxlag_on_instantiation(skg):
    np.random.seed(1312)
    c = np.random.gamma(5, 1, (50, 2))
    np.random.seed(1312)
    v = np.random.weibull(5, 50)
    V = skg.Variogram(c, v, maxlag="median", n_lags=4)
    for b, e in zip([0.88, 1.77, 2.65, 3.53], V.bins):
        assert abs(b - e) < 1e-2


def test_unknown_binning_func(skg, cv30):
    with pytest.raises(ValueError) as e:
        skg.Variogram(*cv30, bin_func="notafunc")
    assert "'notafunc' is not a valid estimator for `bins`" == str(e.value)


def test_invalid_binning_func(skg, cv30):
    with pytest.raises(AttributeError) as e:
        V = skg.Variogram(*cv30)
        V.set_bin_func(42)
    assert "of type string" in str(e.value)


def test_unknown_model(skg, cv30):
    with pytest.raises(ValueError) as e:
        skg.Variogram(*cv30, model="unknown")
    assert ("The theoretical Variogram function unknown is not understood, please provide the "
            "function") == str(e.value)


def test_unsupported_n_lags(skg, cv30):
    with pytest.raises(ValueError) as e:
        skg.Variogram(*cv30, n_lags=15.7)
    assert "n_lags has to be a positive integer" == str(e.value)


def test_value_warning(skg, cv30):
    with pytest.warns(Warning, match="All input values are the same."):
        skg.Variogram(cv30[0], [42] * 30, fit_method="lm")


def test_value_error_trf(skg, cv30):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(AttributeError) as e:
            skg.Variogram(cv30[0], [42] * 30, fit_method="trf")
        assert "'trf' is bounded and therefore" in str(e.value)
        with pytest.raises(AttributeError) as e:
            v = skg.Variogram(cv30[0], [42] * 30, fit_method="lm")
            v.fit_method = "trf"
        assert "'trf' is bounded and therefore" in str(e.value)


def test_pairwise_diffs(skg, cv30):
    V = skg.Variogram(*cv30)
    diff = pdist(np.column_stack((cv30[1], np.zeros(len(cv30[1])))), metric="euclidean")
    assert_array_almost_equal(V.pairwise_diffs, diff, decimal=2)
    V._diff = None
    assert_array_almost_equal(V.pairwise_diffs, diff, decimal=2)


# ---- TestVariogramArguments (:221-702) --------------------------------------------

def test_binning_change_nlags(skg, cv30):
    V = skg.Variogram(*cv30, n_lags=5)
    assert len(V.bins) == 5
    V.n_lags = 8
    assert len(V.bins) == 8


def test_set_bins_directly(skg, cv30):
    V = skg.Variogram(*cv30, n_lags=5)
    bins = [1, 2, 6, 8, 10]
    V.bins = bins
    assert len(V.bins) == 5
    assert_array_almost_equal(V.bins, bins, decimal=8)


def test_binning_callable_arg(skg, cv30):
    def even_func(distances, n, maxlag):
        return np.linspace(0, np.min([np.nanmax(distances), maxlag]), n + 1)[1:], None
    V = skg.Variogram(*cv30, n_lags=8, bin_func=even_func, maxlag=30.0)
    V2 = skg.Variogram(*cv30, n_lags=8, maxlag=30.0)
    assert_array_almost_equal(V.bins, V2.bins, decimal=5)
    assert_array_almost_equal(V.bin_count, V2.bin_count)


def test_binning_iterable_arg(skg, cv30):
    bins = [1, 3, 5, 9, 14, 22]
    V = skg.Variogram(*cv30, bin_func=bins)
    assert V.n_lags == 6 and V.maxlag == 22
    assert_array_almost_equal(V.bins, bins)
    assert V._bin_func_name == "custom_bin_edges"


def test_estimator_method_setting(skg, cv30):
    V = skg.Variogram(*cv30, n_lags=4)
    estimator_list = ("cressie", "matheron", "dowd")
    for i, estimator in enumerate(estimator_list):
        V.set_estimator(estimator)
        assert V.estimator.__name__ == estimator
        assert np.all(np.isfinite(V.experimental))


def test_set_estimator_wrong_type_and_unknown(skg, cv30):
    V = skg.Variogram(*cv30)
    with pytest.raises(ValueError) as e:
        V.set_estimator(45)
    assert str(e.value) == "The estimator has to be a string or callable."
    with pytest.raises(ValueError) as e:
        V.set_estimator("notaestimator")
    assert str(e.value) == ("Variogram estimator notaestimator is not understood, please provide "
                            "the function.")


def test_maxlag_custom_value_and_use_nugget(skg, cv30):
    V = skg.Variogram(*cv30)
    V.maxlag = 33.3
    assert abs(V.maxlag - 33.3) < 0.1
    assert V.use_nugget is False
    V.use_nugget = True
    assert V.use_nugget is True
    with pytest.raises(ValueError) as e:
        V.use_nugget = 42
    assert str(e.value) == "use_nugget has to be of type bool."


def test_n_lags_change_and_exception(skg, cv30):
    V = skg.Variogram(*cv30, n_lags=10)
    assert len(V.bins) == 10
    V.n_lags = 5
    assert len(V.bins) == 5
    for bad in (-5, 15.7):
        with pytest.raises(ValueError) as e:
            skg.Variogram(*cv30, n_lags=bad)
        assert str(e.value) == "n_lags has to be a positive integer"
    with pytest.raises(NotImplementedError):
        skg.Variogram(*cv30, n_lags="auto")


def test_set_values_and_matrices(skg, cv30):
    V = skg.Variogram(*cv30)
    with pytest.raises(ValueError):  # shape mismatch (:542-552)
        V.values = np.arange(5)
    new = np.random.normal(5, 1, 30)
    V.values = new
    assert_array_almost_equal(V.values, new)
    assert V.value_matrix.shape == (30, 30)
    assert V.distance_matrix.shape == (30, 30)
    assert_array_almost_equal(V.distance, pdist(cv30[0]), decimal=12)


def test_nofit(skg, cv30):
    V = skg.Variogram(*cv30, fit_method=None)
    assert V.fit_method is None and V.cov is None and V.cof is None


def test_callable_estimator_on_lag_classes(skg, cv30):
    V = skg.Variogram(*cv30, n_lags=6, estimator=lambda x: float(np.mean(x)) if x.size else np.nan)
    classes = list(V.lag_classes())
    assert_array_almost_equal(V.experimental, [np.mean(k) for k in classes])


# ---- TestVariogramFittingProcedure (:704-1116) ------------------------------------

def test_fit_sigma_variants(fitV):
    V = fitV[0]
    V.fit_sigma = None
    assert V.fit_sigma is None
    sigs = [.8, .5, 2., 2., 5.]
    V.fit_sigma = sigs
    assert list(V.fit_sigma) == sigs
    V.fit_sigma = (0, 1, 2)
    with pytest.raises(AttributeError) as e:
        V.fit_sigma
    assert "len(fit_sigma)" in str(e.value)
    V.fit_sigma = "notAnFunction"
    with pytest.raises(ValueError) as e:
        V.fit_sigma
    assert "fit_sigma is not understood." in str(e.value)


@pytest.mark.parametrize("name,sigma,params,dec", [
    ("linear", [.2, .4, .6, .8, 1.], [13., 0.3, 18.], 8),
    ("exp", [0.0067, 0.0821, 0.1889, 0.2865, 0.3679], [25., 0.2, 18.5], 4),
    ("sqrt", [0.447, 0.632, 0.775, 0.894, 1.], [19.7, 1.5, 16.4], 3),
    ("sq", [0.04, 0.16, 0.36, 0.64, 1.], [5.4, 0.1, 18.5], 2),
])
def test_fit_sigma_functions(fitV, name, sigma, params, dec):
    V = fitV[0]
    V.fit_sigma = name
    assert_array_almost_equal(V.fit_sigma, sigma, decimal=dec)
    V.fit()
    assert_array_almost_equal(V.parameters, params, decimal=1)
    V2 = fitV[0]
    V2.fit(sigma=name)
    assert_array_almost_equal(V2.parameters, params, decimal=1)


def test_fit_lm(skg):
    z = np.load(os.path.join(GOLDEN, "ref_sample.npz"))
    V = skg.Variogram(z["coords"], z["values"], use_nugget=True, n_lags=8, fit_method="lm")
    assert_array_almost_equal(V.parameters, [162.3, 0.5, 0.8], decimal=1)


def test_fitted_model_and_implicit_fit(fitV):
    V = fitV[0]
    V._fit_method = "trf"
    V.fit_sigma = None
    result = np.array([12.48, 17.2, 17.2, 17.2])
    assert_array_almost_equal(result, list(map(V.fitted_model, np.arange(0, 20, 5))), decimal=2)
    V.cof = None
    assert_array_almost_equal(result, list(map(V.fitted_model, np.arange(0, 20, 5))), decimal=2)
    V.cof = None
    assert_array_almost_equal(result, V.transform(np.arange(0, 20, 5)), decimal=2)
    with pytest.raises(AttributeError) as e:
        V.fit(method="unsupported")
    assert "fit_method has to be one of" in str(e.value)


def test_ml_default_and_sq_sigma(skg):
    """:903-932 (reference fixture tests/sample.csv)."""
    z = np.load(os.path.join(GOLDEN, "ref_sample.npz"))
    V = skg.Variogram(z["coords"], z["values"], use_nugget=True, n_lags=15, fit_method="ml")
    assert_array_almost_equal(V.parameters, np.array([41.18, 1.2, 0.]), decimal=2)
    V = skg.Variogram(z["coords"], z["values"], use_nugget=True, n_lags=15, fit_method="ml",
                      fit_sigma="sq")
    assert_array_almost_equal(V.parameters, np.array([42.72, 1.21, 0.]), decimal=2)


def test_manual_fit(skg, fitV):
    _, c, v = fitV
    V = skg.Variogram(c, v, fit_method="manual", model="spherical", fit_range=10., fit_sill=5.)
    assert V.parameters == [10., 5., 0.0]
    V = skg.Variogram(c, v, fit_method="trf", model="matern")
    V._fit_method = "manual"
    V.fit(range=10, sill=5, shape=3)
    assert V.parameters == [10., 5., 3., 0.0]
    with pytest.raises(AttributeError):
        skg.Variogram(c, v, fit_method="manual")
    V = skg.Variogram(c, v, fit_method="trf", n_lags=8)
    params = V.parameters
    V._fit_method = "manual"
    V.fit(sill=14)
    params[1] = 14.
    assert_array_almost_equal(V.parameters, params, decimal=1)
    V = skg.Variogram(c, v, use_nugget=False)
    assert V.parameters[-1] < 1e-10
    V.fit(method="manual", sill=5., nugget=2.)
    assert abs(V.parameters[-1] - 2.) < 1e-10


# ---- TestVariogramMethods (:1241-1446) --------------------------------------------

def test_get_empirical_clone_data(skg, cv30):
    V = skg.Variogram(*cv30, normalize=False, n_lags=10)
    emp_x, emp_y = V.get_empirical()
    assert_array_almost_equal(V.bins, emp_x)
    assert_array_almost_equal(V.experimental, emp_y)
    W = skg.Variogram(*cv30)
    W.bins = [4, 8, 9, 12, 15]
    emp_x, _ = W.get_empirical(bin_center=True)
    assert_array_almost_equal(emp_x, [2., 6., 8.5, 10.5, 13.5])
    copy = V.clone()
    assert_array_almost_equal(copy.experimental, V.experimental)
    assert_array_almost_equal(copy.bins, V.bins)
    lags, var = V.data(n=10, force=False)
    assert_array_almost_equal(lags, [0., 4.7, 9.4, 14.1, 18.8, 23.5, 28.2, 32.9, 37.6, 42.3],
                              decimal=2)
    assert_array_almost_equal(var, [0., 11.82, 13.97, 13.97, 13.97, 13.97, 13.97, 13.97, 13.97,
                                    13.97], decimal=2)


def test_describe_and_parameters(skg, cv30):
    V = skg.Variogram(*cv30, normalize=False, n_lags=10)
    keys = ["normalized_effective_range", "normalized_sill", "normalized_nugget",
            "effective_range", "sill", "nugget"]
    for model_name in ["spherical", "gaussian", "exponential", "cubic", "matern", "stable"]:
        V.set_model(model_name)
        d = V.describe()
        for key in keys:
            assert key in d
        if model_name == "matern":
            assert "smoothness" in d
        if model_name == "stable":
            assert "shape" in d
    V.set_model("matern")
    assert_array_almost_equal(V.parameters, [42.3, 16.2, 0.1, 0.], decimal=2)
    V.set_model("stable")
    assert_array_almost_equal(V.parameters, [42.3, 15.79, 0.45, 0.], decimal=2)
    d = V.describe()
    assert d["params"]["bin_func"] == "even" and d["params"]["n_lags"] == 10
    assert d["estimator"] == "matheron" and d["dist_func"] == "euclidean"


def test_model_quality_measures_match_reference(skg):
    """Variogram.py:2243-2553 (residual statistics, Pearson r, Nash-Sutcliffe, to_DataFrame) against the
    reference's values on its own sample data (tests/golden/ref_quality.json, generated with the
    reference imported from /root/reference)."""
    import json
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_sample.npz"))
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_quality.json")))
    for model, want in ref.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            V = skg.Variogram(g["coords"], g["values"], model=model, n_lags=12, maxlag="median")
            got = dict(mean_residual=V.mean_residual, root_mean_square=V.root_mean_square, rss=V.rss,
                       nrmse=V.nrmse, nrmse_r=V.nrmse_r, r=V.r, NS=V.NS, rmse=V.rmse)
            frame = V.to_DataFrame(n=7)
        for k, val in got.items():
            assert_allclose(val, want[k], rtol=1e-6, err_msg="%s %s" % (model, k))
        assert list(frame.columns) == ["lags", model]
        assert_allclose(frame[model].values, want["frame"], rtol=1e-6)
        exp, mod = V.model_deviations()
        assert exp.shape == mod.shape == (12,)
        assert V.pearson == V.r
    V.update_kwargs(percentile=25)
    assert V._kwargs["percentile"] == 25
