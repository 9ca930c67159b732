"""CPU: host logic of the sparse-space pair filter (PairEngine._filter_thr / _unmap): lag thresholds
clipped at max_dist with a leading, discarded class for coincident points - no device needed."""
import numpy as np
from numpy.testing import assert_array_equal

from skgstat_b200._exact import sq_threshold_bits, sq_upper_inclusive_bits
from skgstat_b200.engine import PairEngine


def _engine(max_dist):
    eng = PairEngine.__new__(PairEngine)          # no CUDA: only the filter state is touched
    eng.pair_filter = None if max_dist is None else (int(sq_upper_inclusive_bits(float(max_dist))), True)
    return eng


def test_no_filter_is_identity():
    thr = sq_threshold_bits(np.array([1.0, 2.0, 3.0]))
    out, src = _engine(None)._filter_thr(thr)
    assert out is thr and src is None
    assert PairEngine._unmap(np.array([5, 6, 7]), None) is not None


def test_thresholds_are_clipped_and_zero_class_is_prepended():
    edges = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    thr = sq_threshold_bits(edges)
    eng = _engine(2.5)
    thr2, src = eng._filter_thr(thr)
    hi = eng.pair_filter[0]
    # class 0 of thr2 = {s == 0}; user classes 0, 1 unchanged; class 2 clipped at max_dist; 3, 4 empty
    assert_array_equal(thr2, np.array([1, thr[0], thr[1], hi], dtype=np.uint64))
    assert_array_equal(src, [1, 2, 3, -1, -1])
    counts2 = np.array([9, 10, 20, 30])             # 9 coincident pairs are dropped
    assert_array_equal(PairEngine._unmap(counts2, src), [10, 20, 30, 0, 0])
    med = PairEngine._unmap(np.array([0.0, 1.5, 2.5, 3.5]), src, fill=np.nan)
    assert_array_equal(np.isnan(med), [False, False, False, True, True])
    # a bound beyond every edge only adds the zero class
    thr3, src3 = _engine(100.0)._filter_thr(thr)
    assert_array_equal(thr3[1:], thr)
    assert_array_equal(src3, [1, 2, 3, 4, 5])
    # a bound below the first edge keeps one clipped class
    thr4, src4 = _engine(0.5)._filter_thr(thr)
    assert thr4.size == 2 and src4[0] == 1 and (src4[1:] == -1).all()
