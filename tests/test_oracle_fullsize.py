"""CPU: the C restatement used for the full-size golden vectors (oracle/pairs_oracle.c through
oracle/fullsize.py) against the dense, literal oracle (scipy pdist + numpy) at a size both run."""
import numpy as np
from numpy.testing import assert_allclose, assert_array_equal

from oracle import fullsize, pairs
from oracle import variogram as ov
from oracle.fields import gaussian_field, random_points


def test_numpy_percentile_restatement():
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 10, 101, 1000):
        a = np.sort(rng.random(n) * 100)
        for q in (0, 10, 33, 50, 73.5, 99, 100):
            got = pairs.percentile_from(n, q, lambda r: a[r])
            assert got == np.percentile(a, q), (n, q)
        assert pairs.median_from(n, lambda r: a[r]) == np.median(a)


def test_isotropic_matches_dense_oracle():
    c = random_points(1300, 1000.0, 2, seed=11)
    v = gaussian_field(c, 100.0, seed=11)
    got = fullsize.isotropic(c, v, ("matheron", "cressie", "dowd"), n_lags=25)
    for est in ("matheron", "cressie", "dowd"):
        ref = ov.dense_variogram(c, v, est, "even", 25, "median")
        assert got["maxlag"] == ref["maxlag"]
        assert_array_equal(got["bins"], ref["bins"])
        assert_array_equal(got["bin_count"], ref["bin_count"])
        assert_allclose(got[est], ref["experimental"], rtol=1e-12, atol=0)


def test_directional_uniform_matches_dense_oracle():
    c = random_points(900, 1000.0, 2, seed=12)
    v = gaussian_field(c, 100.0, seed=12)
    got = fullsize.directional_uniform(c, v, 45, 22.5, n_lags=10)
    ref = ov.dense_directional(c, v, "matheron", "uniform", 10, None, 45, 22.5, "q33", "triangle")
    assert got["n_guard"] == 0
    assert got["bandwidth"] == ref["bandwidth"]
    assert got["mask_count"] == ref["mask_count"]
    assert_array_equal(got["bins"], ref["bins"])
    assert_array_equal(got["bin_count"], ref["bin_count"])
    assert_allclose(got["matheron"], ref["experimental"], rtol=1e-12, atol=0)
