"""Host logic of the histogram-rule binners: numpy's bin-width rules restated on scalars
(skgstat_b200.binning.auto_edges) against np.histogram_bin_edges itself."""
import os
import sys

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                "scikit-gstat_b200"))
from skgstat_b200 import binning  # noqa: E402


def _stats(d):
    n = d.size
    mean = d.sum() / n
    var = (d * d).sum() / n - mean * mean
    std = np.sqrt(var)
    m3 = (d ** 3).sum() / n - 3 * mean * (d * d).sum() / n + 2 * mean ** 3
    srt = np.sort(d)
    q = lambda p: binning.percentile_from_order_stats(n, p, lambda r: srt[r])
    return dict(q25=q(25), q75=q(75), std=std, g1=m3 / std ** 3)


@pytest.mark.parametrize("seed,size", [(0, 50), (1, 777), (2, 5000), (3, 40000)])
@pytest.mark.parametrize("method", binning.AUTO_BINNERS)
def test_auto_edges_equal_numpy(method, seed, size):
    rng = np.random.default_rng(seed)
    d = np.abs(rng.gamma(3.0, 20.0, size)) + rng.random(size)
    st = _stats(d)
    keep = {k: v for k, v in st.items()
            if (k in ("q25", "q75") and method in ("fd", "auto")) or
               (k == "std" and method in ("scott", "doane")) or (k == "g1" and method == "doane")}
    got = binning.auto_edges(method, d.size, d.min(), d.max(), **keep)
    ref = np.histogram_bin_edges(d, bins=method)[1:]
    assert_array_equal(got, ref)
    edges, n = binning.auto_derived_lags(d, method, None)
    assert n == len(ref)
    assert_array_equal(edges, ref)


def test_auto_edges_degenerate_inputs():
    assert_array_equal(binning.auto_edges("sturges", 5, 2.0, 2.0), np.histogram_bin_edges(np.full(5, 2.0), "sturges")[1:])
    a = np.array([1.0, 2.0, 2.0, 2.0, 2.0, 2.0, 3.0])      # IQR = 0: FD width 0 -> one bin
    assert_array_equal(binning.auto_edges("fd", a.size, 1.0, 3.0, q25=np.percentile(a, 25), q75=np.percentile(a, 75)),
                       np.histogram_bin_edges(a, "fd")[1:])
    with pytest.raises(ValueError):
        binning.auto_edges("nope", 10, 0.0, 1.0)
