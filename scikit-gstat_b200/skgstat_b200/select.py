"""Exact order statistics over implicit pair keys (host orchestration).

The device never materialises the N(N-1)/2 keys (squared distances or |dz|).
Exact k-th smallest keys - what np.median / np.percentile / np.nanpercentile /
np.nanmedian need (binning.py:70-79, Variogram.py:1288-1293,
DirectionalVariogram.py:509, estimators.py:178) - are found by sweeps of the
pair kernels:

  1. histogram pass: per segment, count keys per bucket of a monotone bucket
     function over the window [lo, hi)  (skg_pair_select_hist);
  2. locate the bucket holding each wanted rank from the prefix sums;
  3. gather pass: collect the keys of exactly those buckets
     (skg_pair_select_gather), sort the few candidates, pick the ranks.
  Buckets too crowded to gather (ties on lattice data) become the next window
  and the passes repeat; a window that is a single bit pattern is its own answer.

Everything is exact: bucket membership is decided by the same monotone device
function in both passes, counts are integers.

``run_hist`` / ``run_gather`` are callables so the logic is testable on the CPU
against a numpy emulation (tests/test_select_host.py).
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np

from ._exact import bucket_lower_bits, from_bits


def _scale_for(lo_bits: int, hi_bits: int, n_buckets: int) -> float:
    lo, hi = from_bits(lo_bits), from_bits(hi_bits - 1)
    span = hi - lo
    if not (span > 0.0):
        return 0.0
    sc = n_buckets / span
    if not np.isfinite(sc) or not np.isfinite(sc * span):
        return 0.0
    return float(sc)


def exact_select(nseg: int, lo_bits: Sequence[int], hi_bits: Sequence[int], n_buckets: int,
                 rank_fn: Callable[[int, int], List[int]],
                 run_hist: Callable, run_gather: Callable,
                 gather_limit: int = 1 << 24, bucket_limit: int = 1 << 22):
    """Returns (values, totals): values[g] maps every rank of rank_fn(g, totals[g]) to the
    key of that (0-based) rank among the keys of segment g inside its initial window;
    totals[g] = number of keys in that window."""
    values: List[dict] = [dict() for _ in range(nseg)]
    totals = [0] * nseg
    # pending windows: per segment a list of (lo_bits, hi_bits, base, ranks)
    pending: List[list] = [[(int(lo_bits[g]), int(hi_bits[g]), 0, None)] for g in range(nseg)]
    first = True
    while any(pending):
        cur = [p.pop() if p else None for p in pending]
        lo = np.zeros(nseg, dtype=np.uint64)
        hi = np.zeros(nseg, dtype=np.uint64)
        sc = np.zeros(nseg, dtype=np.float64)
        for g, w in enumerate(cur):
            if w is None or w[1] <= w[0]:
                continue
            lo[g], hi[g] = w[0], w[1]
            sc[g] = _scale_for(w[0], w[1], n_buckets)
        hist = np.asarray(run_hist(lo, hi, sc)).reshape(nseg, n_buckets).astype(np.int64)
        # runs to gather: (g, lo_bits, hi_bits, rank of the first key of the run, keys in the run, ranks).
        # A run is the key range of one bucket (or of neighbouring wanted buckets of a segment): the
        # gather pass gets exactly that range as the segment's window (one bucket, scale 0), so its
        # tile culling works on the narrow range instead of the whole histogram window.
        runs = []
        for g, w in enumerate(cur):
            if w is None:
                continue
            w_lo, w_hi, w_base, ranks = w
            row = hist[g]
            tot = int(row.sum())
            if first:
                totals[g] = tot
                ranks = sorted(set(int(r) for r in rank_fn(g, tot)))
                for r in ranks:
                    if not (0 <= r < tot):
                        raise ValueError("rank %d outside [0, %d) in segment %d" % (r, tot, g))
            if not ranks:
                continue
            cum = np.cumsum(row)
            rel = np.asarray(ranks, dtype=np.int64) - w_base
            buckets = np.searchsorted(cum, rel, side="right")
            seg_runs = []
            for b in np.unique(buckets):
                b = int(b)
                rs = [r for r, bb in zip(ranks, buckets) if bb == b]
                below = int(cum[b - 1]) if b > 0 else 0
                cnt = int(row[b])
                if w_hi - w_lo == 1:  # one bit pattern: every key in the window is equal
                    for r in rs:
                        values[g][r] = from_bits(w_lo)
                    continue
                single_bucket = sc[g] == 0.0
                if single_bucket:
                    nlo, nhi = w_lo, w_hi
                else:
                    nlo = bucket_lower_bits(b, w_lo, w_hi, float(sc[g]), n_buckets)
                    nhi = bucket_lower_bits(b + 1, w_lo, w_hi, float(sc[g]), n_buckets) \
                        if b + 1 < n_buckets else w_hi
                if cnt <= bucket_limit or single_bucket:
                    prev = seg_runs[-1] if seg_runs else None
                    if prev is not None and prev[2] == nlo and prev[4] + cnt <= bucket_limit:
                        prev[2], prev[4] = nhi, prev[4] + cnt   # neighbouring bucket: one run
                        prev[5].extend(rs)
                    else:
                        seg_runs.append([g, nlo, nhi, w_base + below, cnt, list(rs)])
                else:
                    pending[g].append((nlo, nhi, w_base + below, rs))
            runs.extend(seg_runs)
        first = False
        # gather calls: at most one run per segment and call, total keys within the capacity limit
        while runs:
            batch, cap, used, rest = [], 0, set(), []
            for run in runs:
                if run[0] in used or (batch and cap + run[4] > gather_limit):
                    rest.append(run)
                else:
                    batch.append(run)
                    used.add(run[0])
                    cap += run[4]
            runs = rest
            glo = np.zeros(nseg, dtype=np.uint64)
            ghi = np.zeros(nseg, dtype=np.uint64)
            gsc = np.zeros(nseg, dtype=np.float64)   # scale 0: every key of the window is bucket 0
            bitmap = np.zeros((nseg * n_buckets + 31) // 32, dtype=np.uint32)
            for g, nlo, nhi, _, _, _ in batch:
                glo[g], ghi[g] = nlo, nhi
                slot = g * n_buckets
                bitmap[slot >> 5] |= np.uint32(1 << (slot & 31))
            if getattr(run_gather, "device_pick", False):
                # the runner sorts the gathered keys where they are and returns only the wanted ranks
                want = [(g * n_buckets, sorted(set(int(r - base) for r in rs)))
                        for g, _, _, base, cnt, rs in batch]
                picked = run_gather(glo, ghi, gsc, bitmap, cap, want)
                for g, _, _, base, cnt, rs in batch:
                    got, vals = picked[g * n_buckets]
                    if got != cnt:
                        raise RuntimeError("segment %d: gathered %d keys, counted %d" % (g, got, cnt))
                    for r in rs:
                        values[g][r] = vals[int(r - base)]
                continue
            keys, slots = run_gather(glo, ghi, gsc, bitmap, cap)
            keys = np.asarray(keys, dtype=np.float64)
            slots = np.asarray(slots, dtype=np.int64)
            if keys.size != cap:
                raise RuntimeError("gather returned %d keys, histogram promised %d" % (keys.size, cap))
            order = np.argsort(slots, kind="stable")
            keys, slots = keys[order], slots[order]
            for g, _, _, base, cnt, rs in batch:
                slot = g * n_buckets
                a = np.searchsorted(slots, slot, side="left")
                z = np.searchsorted(slots, slot, side="right")
                if z - a != cnt:
                    raise RuntimeError("segment %d: gathered %d keys, counted %d" % (g, z - a, cnt))
                idx = sorted(set(int(r - base) for r in rs))
                vals = np.partition(keys[a:z], idx)   # exact order statistics, O(n)
                for r in rs:
                    values[g][r] = float(vals[r - base])
    return values, totals
