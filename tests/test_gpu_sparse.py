"""GPU parity for sparse spaces, MetricSpace(max_dist=...) (MetricSpace.py:141-147,
Variogram.py:1202-1227, 1910-1936) against reference outputs (tests/golden/sparse.npz): only pairs
with 0 < d <= max_dist exist - for maxlag, bins, counts, estimators and the materialised vectors."""
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


def test_sparse_spaces_match_reference(skg, golden_sparse):
    g = golden_sparse
    for case in g.manifest:
        c, v = g.inputs(case["dataset"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ms = skg.MetricSpace(c, max_dist=case["max_dist"])
            V = skg.Variogram(ms, v, estimator=case["estimator"], bin_func=case["bin_func"],
                              n_lags=case["n_lags"], maxlag=case["maxlag"], fit_method=None)
            exp = V.experimental
        ml = g.out(case, "maxlag")
        if np.isnan(ml):
            assert V.maxlag is None
        else:
            assert V.maxlag == ml, str(case)
        assert_array_equal(V.bins, g.out(case, "bins"), err_msg=str(case))
        assert_array_equal(V.bin_count, g.out(case, "bin_count"), err_msg=str(case))
        assert_allclose(exp, g.out(case, "experimental"), rtol=1e-12, equal_nan=True, err_msg=str(case))
        if case["estimator"] == "matheron" and case["bin_func"] == "even" and case["maxlag"] is None:
            assert_array_equal(V.distance, g.out(case, "distance"))
            assert_array_equal(V.pairwise_diffs, g.out(case, "diffs"))
    assert len(g.manifest) == 47


def test_sparse_matrix_accessor_and_limits(skg):
    from scipy import sparse
    from oracle.fields import gaussian_field, random_points
    from oracle import variogram as ov
    c = random_points(500, 100.0, 2, seed=12)
    v = gaussian_field(c, 10.0, seed=12)
    ms = skg.MetricSpace(c, max_dist=25.0)
    m = ms.dists
    assert sparse.issparse(m) and m.shape == (500, 500)
    d = ov.distance(c)
    assert m.nnz == 2 * int(((d > 0) & (d <= 25.0)).sum())
    assert_allclose(m.max(), d[d <= 25.0].max(), rtol=0)
    # a max_dist beyond every distance changes nothing but the zero-distance rule
    dense = skg.Variogram(c, v, n_lags=9, fit_method=None)
    wide = skg.Variogram(skg.MetricSpace(c, max_dist=1e6), v, n_lags=9, fit_method=None)
    assert_array_equal(dense.bins, wide.bins)
    assert_array_equal(dense.bin_count, wide.bin_count)
    assert_allclose(dense.experimental, wide.experimental, rtol=1e-13)
    # batched sweep and cross-variograms honour the filter too
    V = skg.Variogram(ms, v, n_lags=8, fit_method=None)
    draws = np.vstack([v + 0.1 * k for k in range(3)])
    got = V.experimental_batch(draws)
    for k in range(3):
        one = skg.Variogram(ms, draws[k], n_lags=8, fit_method=None)
        assert_allclose(got[k], one.experimental, rtol=1e-13)
