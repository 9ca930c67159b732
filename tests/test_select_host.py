"""CPU: the exact order-statistic orchestration (skgstat_b200/select.py) against a
numpy emulation of the two device passes, including tie-heavy lattice keys that
force window refinement."""
import numpy as np
import pytest

from skgstat_b200._exact import from_bits, sq_threshold, sq_upper_inclusive_bits, to_bits
from skgstat_b200.select import exact_select


def emulate(keys_per_seg, n_buckets):
    def buckets(g, lo, hi, sc):
        k = keys_per_seg[g]
        kb = k.view(np.uint64)
        m = (kb >= lo[g]) & (kb < hi[g])
        lo_val = np.array([lo[g]], dtype=np.uint64).view(np.float64)[0]
        t = (k[m] - lo_val) * sc[g]
        b = np.minimum(n_buckets - 1, t.astype(np.int64))
        return k[m], b

    def run_hist(lo, hi, sc):
        out = np.zeros((len(keys_per_seg), n_buckets), dtype=np.uint64)
        for g in range(len(keys_per_seg)):
            _, b = buckets(g, lo, hi, sc)
            out[g] = np.bincount(b, minlength=n_buckets)
        return out

    def run_gather(lo, hi, sc, bitmap, cap):
        ks, ss = [], []
        for g in range(len(keys_per_seg)):
            k, b = buckets(g, lo, hi, sc)
            slot = g * n_buckets + b
            hit = ((bitmap[slot >> 5] >> (slot & 31).astype(np.uint32)) & 1).astype(bool)
            ks.append(k[hit])
            ss.append(slot[hit])
        return np.concatenate(ks), np.concatenate(ss)

    return run_hist, run_gather


@pytest.mark.parametrize("n_buckets,bucket_limit", [(64, 1 << 22), (16, 8), (1024, 3)])
def test_exact_select_matches_sort(n_buckets, bucket_limit):
    rng = np.random.default_rng(3)
    segs = [rng.random(5000) * 37.0,
            np.floor(rng.random(4000) * 12.0),          # heavy ties
            np.abs(rng.normal(size=3000)) * 1e-3,
            np.zeros(0)]
    hi = [to_bits(float(s.max())) + 1 if s.size else 1 for s in segs]
    lo = [0] * len(segs)

    def rank_fn(g, tot):
        if tot == 0:
            return []
        return sorted({0, tot // 3, (tot - 1) // 2, tot // 2, tot - 1})

    rh, rg = emulate(segs, n_buckets)
    vals, totals = exact_select(len(segs), lo, hi, n_buckets, rank_fn, rh, rg,
                                bucket_limit=bucket_limit, gather_limit=1 << 12)
    for g, s in enumerate(segs):
        assert totals[g] == s.size
        srt = np.sort(s)
        for r in rank_fn(g, s.size):
            assert vals[g][r] == srt[r], (g, r)


def test_window_restricts_keys():
    rng = np.random.default_rng(5)
    d = rng.random(20000) * 50
    s = d * d
    limit = 20.0
    hi = sq_upper_inclusive_bits(limit)
    rh, rg = emulate([s], 256)
    inside = np.sort(s[np.sqrt(s) <= limit])
    vals, totals = exact_select(1, [0], [hi], 256, lambda g, t: [0, t // 2, t - 1], rh, rg)
    assert totals[0] == inside.size
    assert vals[0][inside.size // 2] == inside[inside.size // 2]
    assert vals[0][inside.size - 1] == inside[-1]


def test_thresholds_reproduce_digitize():
    rng = np.random.default_rng(0)
    s = rng.random(100000) * 1e4
    d = np.sqrt(s)
    edges = np.sort(rng.random(12) * 100)
    T = np.array([sq_threshold(e) for e in edges])
    assert (np.digitize(d, edges) == np.searchsorted(T, s, side="right")).all()
    x = np.arange(0, 60)
    s = (x[:, None] ** 2 + x[None, :] ** 2).ravel().astype(float)
    d = np.sqrt(s)
    edges = np.linspace(0, d.max(), 11)[1:]
    T = np.array([sq_threshold(e) for e in edges])
    assert (np.digitize(d, edges) == np.searchsorted(T, s, side="right")).all()
    assert from_bits(to_bits(3.25)) == 3.25


def test_single_bucket_with_known_counts_needs_no_histogram_pass():
    """Direct mode of engine.sq_order_stats: one bucket per window, counts known from the cell
    sweep.  The histogram callback is answered from those counts; with the bucket limit lifted
    the select goes straight to the gather (and never loops on a refinement it cannot serve),
    also through the device-pick protocol of the gather callback."""
    rng = np.random.default_rng(5)
    keys = np.sort(rng.random(6000) * 3.0 + 1.0)
    lo_b, hi_b = to_bits(1.5), to_bits(2.5)
    inside = keys[(keys.view(np.uint64) >= lo_b) & (keys.view(np.uint64) < hi_b)]
    _, rg = emulate([keys], 1)
    calls = []

    def rh(lo, hi, sc):
        assert not calls, "a second histogram request means a refinement loop"
        calls.append(1)
        return np.array([[inside.size]], dtype=np.int64)

    ranks = [0, inside.size // 2, inside.size - 1]
    vals, totals = exact_select(1, [lo_b], [hi_b], 1, lambda g, t: ranks, rh, rg,
                                gather_limit=1 << 24, bucket_limit=1 << 62)
    assert totals == [inside.size]
    assert [vals[0][r] for r in ranks] == [inside[r] for r in ranks]

    def rg_pick(lo, hi, sc, bitmap, cap, want=None):   # same protocol as PairEngine's runner
        k, s = rg(lo, hi, sc, bitmap, cap)
        out = {}
        for slot, rel in want:
            srt = np.sort(k[s == slot])
            out[slot] = (srt.size, {r: float(srt[r]) for r in rel})
        return out

    rg_pick.device_pick = True
    calls.clear()
    vals2, _ = exact_select(1, [lo_b], [hi_b], 1, lambda g, t: ranks, rh, rg_pick,
                            gather_limit=1 << 24, bucket_limit=1 << 62)
    assert vals2[0] == vals[0]
    # with the default bucket limit a crowded single bucket asks for a refinement: the guard trips
    calls.clear()
    with pytest.raises(AssertionError):
        exact_select(1, [lo_b], [hi_b], 1, lambda g, t: ranks, rh, rg, bucket_limit=100)


def _bisect_bucket_lower_bits(b, lo_bits, hi_bits, scale, n_buckets):
    """Plain bisection over the bit patterns: the definition bucket_lower_bits has to meet."""
    from skgstat_b200._exact import bucket_of
    lo_val = from_bits(lo_bits)
    a, z = lo_bits, hi_bits
    while a < z:
        mid = (a + z) // 2
        if bucket_of(from_bits(mid), lo_val, scale, n_buckets) >= b:
            z = mid
        else:
            a = mid + 1
    return a


def test_bucket_lower_bits_equals_plain_bisection():
    """The bracketed search (analytic start, doubling steps) returns exactly what the ~62-step
    bisection returns: tiny and wide windows, windows starting at 0, every bucket count used."""
    from skgstat_b200._exact import bucket_lower_bits
    from skgstat_b200.select import _scale_for
    rng = np.random.default_rng(11)
    checked = 0
    for _ in range(600):
        lo = float(rng.choice([0.0, rng.random() * 10, rng.random() * 1e-3, rng.random() * 1e5]))
        width = float(rng.choice([1e-12, 1e-6, 1e-3, 1.0, 100.0])) * (rng.random() + 0.01) * max(lo, 1e-3)
        lb, hb = to_bits(lo), to_bits(lo + width) + 1
        if hb - lb < 2:
            continue
        nb = int(rng.choice([1, 2, 256, 4096, 65536, 1 << 20]))
        sc = _scale_for(lb, hb, nb)
        for b in (0, 1, nb // 2, nb - 1, nb, int(rng.integers(0, nb + 1))):
            if b == 0 and sc == float("inf"):
                continue
            assert bucket_lower_bits(b, lb, hb, sc, nb) == _bisect_bucket_lower_bits(b, lb, hb, sc, nb)
            checked += 1
    assert checked > 2000


def test_merged_dz_windows_cover_their_members():
    """engine._merge_dz_windows: every class keeps a window that contains its own, runs share one
    window, the union is at most WINDOW_MERGE_GROWTH times the widest member, full-range windows
    are left alone."""
    from skgstat_b200.engine import PairEngine
    full = to_bits(9.0) + 1
    lo_v = np.array([0.20, 0.46, 0.73, 0.0, 1.38, 1.42, 1.43, 1.41, 1.45, 1.47, 0.0, 1.40])
    hi_v = np.array([0.23, 0.51, 0.81, 9.0, 1.53, 1.58, 1.60, 1.58, 1.62, 1.65, 9.0, 1.57])
    w_lo = np.array([to_bits(x) for x in lo_v], dtype=np.uint64)
    w_hi = np.array([to_bits(x) + 1 for x in hi_v], dtype=np.uint64)
    w_hi[[3, 10]] = full
    m_lo, m_hi = PairEngine._merge_dz_windows(w_lo, w_hi, full)
    assert np.all(m_lo <= w_lo) and np.all(m_hi >= w_hi)
    assert (m_lo[3], m_hi[3]) == (0, full) and (m_lo[10], m_hi[10]) == (0, full)
    assert len({(int(a), int(b)) for a, b in zip(m_lo[4:10], m_hi[4:10])}) == 1      # one run
    assert (int(m_lo[0]), int(m_hi[0])) == (int(w_lo[0]), int(w_hi[0]))             # distinct medians stay apart
    widths = hi_v - lo_v
    for g in (4, 5, 6, 7, 8, 9):
        assert from_bits(int(m_hi[g])) - from_bits(int(m_lo[g])) <= PairEngine.WINDOW_MERGE_GROWTH * widths[4:10].max() + 1e-12
    assert (int(m_lo[11]), int(m_hi[11])) == (int(w_lo[11]), int(w_hi[11]))         # a full-range class breaks the run
