"""Oracle (test infrastructure, not product): ctypes binding of oracle/pairs_oracle.c plus the
exact order-statistic selection built on its bit-pattern histograms.  Used to pin the
full-size BASELINE.json configurations (oracle/make_fullsize_golden.py) where the reference's
O(N^2) arrays do not fit.  numpy restatements cited inline:

* np.median            -> mean of the two middle order statistics        (Variogram.py:1288-1289)
* np.percentile /      -> 'linear' method: virtual index (n-1)*q/100, numpy's _lerp
  np.nanpercentile        (binning.py:70-79, DirectionalVariogram.py:509)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libpairs_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", _HERE], check=True)
        L = C.CDLL(path)
        p, i64, u64 = C.c_void_p, C.c_int64, C.c_uint64
        L.oracle_key_sweep.argtypes = [p, p, p, i64, p, p, C.c_int, C.c_int, C.c_int, p, p, i64, C.c_int, p, p,
                                       p, p, i64, p, p, p, i64, p, p, i64, i64]
        L.oracle_lag_hist.argtypes = [p, p, p, i64, p, p, C.c_int, p, p, p, p, p, i64, i64]
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


THREADS = int(os.environ.get("ORACLE_THREADS", str(os.cpu_count() or 1)))


def _parallel(fn):
    """fn(row_begin, row_step) on THREADS Python threads (ctypes releases the GIL); list of results."""
    from concurrent.futures import ThreadPoolExecutor
    t = max(1, THREADS)
    if t == 1:
        return [fn(0, 1)]
    with ThreadPoolExecutor(t) as ex:
        return list(ex.map(lambda r: fn(r, t), range(t)))


INF_BITS = 0x7FF0000000000000


class Pairs:
    """All pairs (i < j, caller's order) of one 2-D point set, optionally behind the reference's
    directional mask: dirp = (model 1 triangle / 2 compass, np.radians(azimuth),
    np.radians(tolerance / 2), bandwidth / 2)."""

    def __init__(self, coords, values=None, dirp=None):
        c = np.ascontiguousarray(np.asarray(coords, dtype=np.float64))
        self.x = np.ascontiguousarray(c[:, 0])
        self.y = np.ascontiguousarray(c[:, 1])
        self.z = None if values is None else np.ascontiguousarray(np.asarray(values, dtype=np.float64))
        self.n = len(self.x)
        self.dirp = None if dirp is None else np.ascontiguousarray(np.asarray(dirp, dtype=np.float64))
        self.n_guard = 0

    # ------------------------------------------------------------ order statistics
    NB1, NB2 = 1 << 12, 1 << 12

    def _sweep(self, edges, key_kind, mode, lo1, sc1, targets=None, cap=0):
        nl = 0 if edges is None else len(edges)
        nseg = max(nl, 1)
        e = None if edges is None else np.ascontiguousarray(edges, dtype=np.float64)
        lo1 = np.ascontiguousarray(lo1, dtype=np.float64)
        sc1 = np.ascontiguousarray(sc1, dtype=np.float64)
        K = 0 if targets is None else len(targets["seg"])
        t = {k: np.ascontiguousarray(v) for k, v in (targets or {}).items()}
        L = lib()

        def part(r0, step):
            hist = np.zeros((nseg * self.NB1) if mode == 0 else max(K * self.NB2, 1), dtype=np.uint64)
            out = np.empty(max(K * cap, 1), dtype=np.float64)
            cnt = np.zeros(max(K, 1), dtype=np.int64)
            guard = C.c_uint64(0)
            L.oracle_key_sweep(_ptr(self.x), _ptr(self.y), _ptr(self.z), self.n, _ptr(self.dirp), _ptr(e), nl,
                               key_kind, mode, _ptr(lo1), _ptr(sc1), self.NB1, K,
                               _ptr(t.get("seg")), _ptr(t.get("want1")), _ptr(t.get("lo2")), _ptr(t.get("sc2")),
                               self.NB2, _ptr(t.get("want2")), _ptr(hist), _ptr(out), cap, _ptr(cnt),
                               C.byref(guard), r0, step)
            return hist, out, cnt, int(guard.value)

        parts = _parallel(part)
        if mode == 0:
            self.n_guard += sum(p[3] for p in parts)
            return sum(p[0] for p in parts).reshape(nseg, self.NB1).astype(np.int64)
        if mode == 1:
            return sum(p[0] for p in parts).reshape(K, self.NB2).astype(np.int64)
        keys = []
        for k in range(K):
            got = [p[1][k * cap: k * cap + min(int(p[2][k]), cap)] for p in parts]
            if any(int(p[2][k]) > cap for p in parts):
                raise RuntimeError("collect overflow")
            keys.append(np.sort(np.concatenate(got)))
        return keys

    def order_stats(self, rank_fn, edges=None, key_kind=0, key_max=None):
        """Exact order statistics of the keys (d or |dz|) per lag class: rank_fn(class, total) ->
        0-based ranks.  Returns (list of dict rank -> key, totals).  key_max: an upper bound of the
        keys (default: bounding-box diagonal / value range)."""
        nl = 0 if edges is None else len(edges)
        nseg = max(nl, 1)
        if key_max is None:
            if key_kind == 0:
                key_max = float(np.hypot(np.ptp(self.x), np.ptp(self.y)))
            else:
                key_max = float(np.nanmax(self.z) - np.nanmin(self.z))
        key_max = key_max * (1.0 + 1e-9) + 1e-300
        lo1 = np.zeros(nseg)
        sc1 = np.full(nseg, self.NB1 / key_max)
        h1 = self._sweep(edges, key_kind, 0, lo1, sc1)
        totals = [int(h1[g].sum()) for g in range(nseg)]
        vals = [dict() for _ in range(nseg)]
        tg = dict(seg=[], want1=[], lo2=[], sc2=[], want2=[])
        rel1 = []   # per level-1 target: [(rank, rank inside the bucket)]
        for g in range(nseg):
            cum = np.cumsum(h1[g])
            by = {}
            for r in sorted(set(int(r) for r in rank_fn(g, totals[g]))):
                if not (0 <= r < totals[g]):
                    raise ValueError("rank outside the class")
                b = int(np.searchsorted(cum, r, side="right"))
                by.setdefault(b, []).append((r, r - (int(cum[b - 1]) if b else 0)))
            for b, lst in by.items():
                tg["seg"].append(g); tg["want1"].append(b)
                tg["lo2"].append(lo1[g] + b / sc1[g]); tg["sc2"].append(sc1[g] * self.NB2)
                tg["want2"].append(-1)
                rel1.append(lst)
        if not rel1:
            return vals, totals
        arr = lambda d: dict(seg=np.array(d["seg"], dtype=np.int32), want1=np.array(d["want1"], dtype=np.int64),
                             lo2=np.array(d["lo2"], dtype=np.float64), sc2=np.array(d["sc2"], dtype=np.float64),
                             want2=np.array(d["want2"], dtype=np.int64))
        h2 = self._sweep(edges, key_kind, 1, lo1, sc1, arr(tg))
        t3 = dict(seg=[], want1=[], lo2=[], sc2=[], want2=[])
        rel2, cap = [], 1
        for k, lst in enumerate(rel1):
            if int(h2[k].sum()) != int(h1[tg["seg"][k]][tg["want1"][k]]):
                raise RuntimeError("level-2 histogram does not add up to its level-1 bucket")
            cum = np.cumsum(h2[k])
            by = {}
            for r, rr in lst:
                b = int(np.searchsorted(cum, rr, side="right"))
                by.setdefault(b, []).append((r, rr - (int(cum[b - 1]) if b else 0)))
            for b, l2 in by.items():
                for key in ("seg", "want1", "lo2", "sc2"):
                    t3[key].append(tg[key][k])
                t3["want2"].append(b)
                rel2.append((l2, int(h2[k][b])))
                cap = max(cap, int(h2[k][b]))
        keys = self._sweep(edges, key_kind, 2, lo1, sc1, arr(t3), cap=cap)
        for k, (l2, expect) in enumerate(rel2):
            if keys[k].size != expect:
                raise RuntimeError("collect / histogram mismatch")
            for r, rr in l2:
                vals[t3["seg"][k]][r] = float(keys[k][rr])
        return vals, totals

    # ------------------------------------------------------------------ sweeps
    def lag_hist(self, edges):
        nl = len(edges)
        e = np.ascontiguousarray(edges, dtype=np.float64)
        L = lib()

        def part(r0, step):
            count = np.zeros(nl, dtype=np.uint64)
            ssq = np.zeros(nl, dtype=np.longdouble)
            srt = np.zeros(nl, dtype=np.longdouble)
            guard, masked = C.c_uint64(0), C.c_uint64(0)
            L.oracle_lag_hist(_ptr(self.x), _ptr(self.y), _ptr(self.z), self.n, _ptr(self.dirp), _ptr(e), nl,
                              _ptr(count), _ptr(ssq), _ptr(srt), C.byref(masked), C.byref(guard), r0, step)
            return count, ssq, srt, int(masked.value), int(guard.value)

        parts = _parallel(part)
        self.n_guard += sum(p[4] for p in parts)
        count = sum(p[0] for p in parts).astype(np.int64)
        ssq = np.asarray(sum(p[1] for p in parts), dtype=np.longdouble).astype(np.float64)
        srt = np.asarray(sum(p[2] for p in parts), dtype=np.longdouble).astype(np.float64)
        return count, ssq, srt, sum(p[3] for p in parts)


# numpy restatements on order statistics -----------------------------------------------------

def median_ranks(t):
    return [] if t == 0 else [(t - 1) // 2, t // 2]


def median_from(t, get):
    """np.median: mean of the two middle elements (numpy _median -> np.mean of the slice)."""
    a, b = get((t - 1) // 2), get(t // 2)
    return a if t % 2 else (a + b) / 2.0      # (a + b) / 2 rounds like np.mean([a, b])


def percentile_ranks(t, q):
    """numpy 'linear' method: virtual index (t - 1) * q / 100 (numpy function_base._compute_virtual_index)."""
    virtual = (t - 1) * (q / 100.0)
    lo = int(np.floor(virtual))
    return [min(max(lo, 0), t - 1), min(max(lo + 1, 0), t - 1)], virtual - np.floor(virtual)


def percentile_from(t, q, get):
    """numpy function_base._lerp(a, b, t): a + (b - a) * t, and b - (b - a) * (1 - t) for t >= 0.5."""
    (i0, i1), g = percentile_ranks(t, q)
    a, b = np.float64(get(i0)), np.float64(get(i1))
    g = np.float64(g)
    diff = b - a
    out = a + diff * g
    if g >= 0.5:
        out = b - diff * (1 - g)
    if g == 0:      # numpy: lerp_interpolation where t == 0 -> a  (and b where t == 1)
        out = a
    return float(out)
