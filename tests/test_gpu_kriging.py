"""GPU parity tests for the batched ordinary-kriging kernel."""
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import kriging as okr
from oracle.fields import gaussian_field, random_points

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def skg():
    import skgstat_b200
    skgstat_b200._lib.load()
    return skgstat_b200


def _descr(case):
    d = dict(model=case["model"], effective_range=case["effective_range"], sill=case["sill"],
             nugget=case["nugget"], dist_func="euclidean")
    if case.get("shape") is not None:
        d["shape"] = case["shape"]
    return d


def test_kriging_reference_golden(skg, golden_kriging):
    g = golden_kriging
    for case in g.manifest:
        c, v = g.inputs(case["dataset"])
        t = g.out(case, "targets")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ok = skg.OrdinaryKriging(_descr(case), min_points=case["min_points"],
                                     max_points=case["max_points"], coordinates=c, values=v)
            z = ok.transform(*[t[:, k] for k in range(t.shape[1])])
        zr, sr = g.out(case, "z"), g.out(case, "sigma")
        assert_array_equal(np.isnan(z), np.isnan(zr), err_msg=str(case))
        assert_array_equal(np.isnan(ok.sigma), np.isnan(sr), err_msg=str(case))
        assert ok.no_points_error == case["no_points_error"]
        assert_allclose(z, zr, rtol=1e-9, equal_nan=True, err_msg=str(case))
        assert_allclose(ok.sigma, sr, rtol=1e-9, atol=1e-9 * case["sill"], equal_nan=True,
                        err_msg=str(case))


def test_kriging_from_variogram_object(skg, golden_kriging):
    """End to end through the classes: Variogram (GPU) -> fit (CPU) -> OrdinaryKriging (GPU);
    BASELINE.md section 3 known answers (fit differences enter at ~1e-6)."""
    g = golden_kriging
    c, v = g.inputs("pancake300")
    V = skg.Variogram(c, v, model="exponential", n_lags=25, maxlag="median")
    ok = skg.OrdinaryKriging(V, min_points=5, max_points=64)
    gx, gy = np.mgrid[0:499:6j, 0:499:6j].reshape(2, -1)
    z = ok.transform(gx, gy)
    assert_allclose(z[:4], [216.82668695078203, 213.25019510789258, 208.75398093840414,
                            204.76546428916612], rtol=1e-5)
    assert_allclose(ok.sigma[:4], [821.5990160593558, 607.4358198357556, 457.6864840412478,
                                   250.3293362582594], rtol=1e-5)
    assert np.isnan(z).sum() == 0


@pytest.mark.parametrize("model,shape", [("exponential", None), ("spherical", None),
                                         ("cubic", None), ("stable", 1.4)])
@pytest.mark.parametrize("maxp", [1, 7, 33, 64])
def test_kriging_against_oracle(skg, model, shape, maxp):
    c = random_points(3000, 1000.0, 2, seed=8)
    v = gaussian_field(c, 100.0, seed=8)
    rng = np.random.default_rng(maxp)
    t = rng.random((400, 2)) * 1400.0 - 200.0  # some targets outside the hull / range
    t[:20] = c[:20]                            # targets on top of samples
    rad = 120.0
    minp = min(3, maxp)
    zr, sr, st = okr.krige(c, v, t, model, rad, 4.0, 0.3, shape, minp, maxp)
    d = dict(model=model, effective_range=rad, sill=4.0, nugget=0.3, dist_func="euclidean")
    if shape is not None:
        d["shape"] = shape
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ok = skg.OrdinaryKriging(d, min_points=minp, max_points=maxp, coordinates=c, values=v)
        z = ok.transform(t[:, 0], t[:, 1])
    assert_array_equal(np.isnan(z), np.isnan(zr))
    assert ok.no_points_error == int((st == 1).sum())
    assert_allclose(z, zr, rtol=1e-9, equal_nan=True)
    assert_allclose(ok.sigma, sr, rtol=1e-9, atol=4e-9, equal_nan=True)


@pytest.mark.parametrize("maxp,rad", [(65, 200.0), (100, 250.0), (128, 300.0)])
def test_kriging_wide_neighbourhoods_against_oracle(skg, maxp, rad):
    """max_points beyond 64: the 4-slots-per-lane variants of both solvers.  The
    reference has no limit (Kriging.py:145-158, 305-330); 128 is this library's."""
    c = random_points(2500, 1000.0, 2, seed=18)
    v = gaussian_field(c, 100.0, seed=18)
    rng = np.random.default_rng(maxp)
    t = rng.random((150, 2)) * 1200.0 - 100.0
    t[:10] = c[:10]
    zr, sr, st = okr.krige(c, v, t, "exponential", rad, 4.0, 0.3, None, 5, maxp)
    d = dict(model="exponential", effective_range=rad, sill=4.0, nugget=0.3, dist_func="euclidean")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ok = skg.OrdinaryKriging(d, min_points=5, max_points=maxp, coordinates=c, values=v)
        z = ok.transform(t[:, 0], t[:, 1])
    assert_array_equal(np.isnan(z), np.isnan(zr))
    assert_allclose(z, zr, rtol=1e-9, equal_nan=True)
    assert_allclose(ok.sigma, sr, rtol=1e-9, atol=4e-9, equal_nan=True)
    with pytest.raises(NotImplementedError):
        skg.OrdinaryKriging(d, min_points=5, max_points=129, coordinates=c, values=v).transform(t[:, 0], t[:, 1])


def test_kriging_ties_on_lattice_and_duplicates(skg):
    gx, gy = np.meshgrid(np.arange(30.0), np.arange(30.0))
    c = np.column_stack((gx.ravel(), gy.ravel()))
    c = np.vstack((c, c[:40]))  # duplicated samples are dropped first (Kriging.py:285-303)
    v = gaussian_field(c, 6.0, seed=2)
    v[900:] = v[:40] + 1.0      # the duplicates carry different values; the first one wins
    tx, ty = np.meshgrid(np.arange(-2, 32, 0.5), np.arange(-2, 32, 0.5))
    t = np.column_stack((tx.ravel(), ty.ravel()))
    zr, sr, st = okr.krige(c, v, t, "spherical", 6.0, 2.0, 0.0, None, 4, 9)
    ok = skg.OrdinaryKriging(dict(model="spherical", effective_range=6.0, sill=2.0, nugget=0,
                                  dist_func="euclidean"), min_points=4, max_points=9,
                             coordinates=c, values=v)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        z = ok.transform(t[:, 0], t[:, 1])
    assert_array_equal(np.isnan(z), np.isnan(zr))
    assert_allclose(z, zr, rtol=1e-9, atol=1e-12, equal_nan=True)
    assert_allclose(ok.sigma, sr, rtol=1e-9, atol=2e-9, equal_nan=True)


def test_kriging_argument_validation(skg, golden_kriging):
    """skgstat/tests/test_kriging.py:33-124 (messages asserted by the reference)."""
    c, v = golden_kriging.inputs("pancake300")
    d = dict(model="spherical", effective_range=300.0, sill=1500.0, nugget=0, dist_func="euclidean")
    with pytest.raises(ValueError, match="min_points has to be an integer"):
        skg.OrdinaryKriging(d, min_points=4.0, coordinates=c, values=v)
    with pytest.raises(ValueError, match="min_points can't be negative"):
        skg.OrdinaryKriging(d, min_points=-2, coordinates=c, values=v)
    with pytest.raises(ValueError, match="min_points can't be larger than max_points"):
        skg.OrdinaryKriging(d, min_points=20, max_points=10, coordinates=c, values=v)
    with pytest.raises(ValueError, match="max_points has to be an integer"):
        skg.OrdinaryKriging(d, max_points=16.0, coordinates=c, values=v)
    with pytest.raises(ValueError, match="mode has to be one of"):
        skg.OrdinaryKriging(d, mode="foo", coordinates=c, values=v)
    with pytest.raises(AttributeError, match="solver has to be"):
        skg.OrdinaryKriging(d, solver="peter", coordinates=c, values=v)
