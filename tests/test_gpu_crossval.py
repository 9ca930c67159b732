"""GPU: leave-one-out cross validation - one batched
skg_krige call with exclude[t] = t - against the reference's jacknife."""
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import kriging as okr
from oracle.fields import gaussian_field, random_points

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def skg():
    import skgstat_b200
    skgstat_b200._lib.load()
    return skgstat_b200


def test_leave_one_out_reference_golden(skg, golden_cv):
    from skgstat_b200.util.cross_validation import leave_one_out
    g = golden_cv
    for case in g.manifest:
        c, v = g.inputs(case["dataset"])
        d = dict(model=case["model"], effective_range=case["effective_range"], sill=case["sill"],
                 nugget=case["nugget"], dist_func="euclidean")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dev = leave_one_out(d, g.out(case, "indices"), coordinates=c, values=v)
        ref = g.out(case, "deviations")
        assert_allclose(dev, ref, rtol=1e-8, atol=1e-9 * max(1.0, np.abs(v).max()), err_msg=str(case))
        # metrics as the reference computes them (util/cross_validation.py:66-74)
        rm = g.out(case, "metrics")
        assert_allclose(np.sqrt(np.nanmean(dev ** 2)), rm[0], rtol=1e-8)
        assert_allclose(np.nanmean(dev ** 2), rm[1], rtol=1e-8)
        assert_allclose(np.nansum(np.abs(dev)) / len(dev), rm[2], rtol=1e-8)


def test_cross_validate_method_end_to_end(skg, golden_cv):
    """Variogram.cross_validate through the classes (GPU variogram, CPU fit, GPU jackknife):
    the fit enters at ~1e-6, so compare to the reference metric at 1e-4."""
    g = golden_cv
    case = g.manifest[0]
    c, v = g.inputs(case["dataset"])
    V = skg.Variogram(c, v, model=case["model"], n_lags=15, maxlag="median")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k, metric in enumerate(("rmse", "mse", "mae")):
            got = V.cross_validate(n=case["n"], metric=metric, seed=case["seed"])
            assert_allclose(got, g.out(case, "metrics")[k], rtol=1e-4)
    with pytest.raises(ValueError):
        V.cross_validate(metric="foo")
    with pytest.raises(AttributeError):
        V.cross_validate(method="kfold")


def test_leave_one_out_against_oracle_3d(skg):
    from skgstat_b200.util.cross_validation import leave_one_out
    c = random_points(500, 60.0, 3, seed=6)
    v = gaussian_field(c, 12.0, seed=6)
    d = dict(model="exponential", effective_range=25.0, sill=4.0, nugget=0.2, dist_func="euclidean")
    idx = np.arange(0, 500, 3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dev = leave_one_out(d, idx, min_points=4, max_points=20, coordinates=c, values=v)
    ref = okr.jacknife_deviations(c, v, idx, "exponential", 25.0, 4.0, 0.2, None, 4, 20)
    assert_allclose(dev, ref, rtol=1e-9, atol=1e-10, equal_nan=True)


def test_leave_one_out_with_colocated_samples(skg):
    """duplicated coordinates - the reference removes the point, then de-duplicates
    (util/cross_validation.py:9-21); the drop-in follows it literally for the affected targets."""
    from oracle import kriging as okr
    from skgstat_b200.util import cross_validation as cv
    rng = np.random.default_rng(4)
    c = rng.random((120, 2)) * 100
    v = rng.normal(size=120)
    c[7] = c[3]          # twins with different observations
    c[50] = c[3]
    c[90] = c[61]
    V = skg.Variogram(c, v, model="exponential", n_lags=10, maxlag="median")
    idx = np.array([0, 3, 7, 50, 61, 90, 100])
    got = cv.leave_one_out(V, idx)
    d = V.describe()
    for k, i in enumerate(idx):
        cc, vv = np.delete(c, i, axis=0), np.delete(v, i)
        _, first = np.unique(cc, axis=0, return_index=True)
        first.sort()
        zr, _, _ = okr.krige(cc[first], vv[first], c[i][None, :], "exponential", d["effective_range"],
                             d["sill"], d["nugget"], None, 5, 15)
        np.testing.assert_allclose(got[k], zr[0] - v[i], rtol=1e-8, atol=1e-10)
