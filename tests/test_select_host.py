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
