"""Distance order statistics: sub-sample windows + bulk counts against the log-cell sweep
(developer tool; both are exact, so the values must be identical)."""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import torch
from skgstat_b200 import binning
from skgstat_b200.data import gaussian_field, random_points
from skgstat_b200.engine import PairEngine
warnings.simplefilter("ignore")


def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, r


for n in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["20000", "100000", "200000"])]:
    L = 10000.0 * np.sqrt(n / 200000.0)
    c = random_points(n, L, 2, seed=0)
    eng = PairEngine(c, None)
    med = lambda t: [] if t == 0 else [(t - 1) // 2, t // 2]
    qs = [(i / 10) * 100 for i in range(1, 11)]

    def uni(t):
        out = set()
        for q in qs:
            out.update(binning.percentile_ranks(t, q)[:2])
        return sorted(out)

    for name, fn, lim in (("median", med, None), ("uniform10", uni, None), ("uniform10<=0.3L", uni, 0.3 * L)):
        eng._select_mode = "cells"
        l0 = eng.launches
        t_old, (v_old, tot_old) = timed(lambda: eng.sq_order_stats(fn, limit=lim))
        k_old = (eng.launches - l0) // 4
        eng._select_mode = "auto"
        l0 = eng.launches
        t_new, (v_new, tot_new) = timed(lambda: eng.sq_order_stats(fn, limit=lim))
        k_new = (eng.launches - l0) // 4
        same = tot_old == tot_new and v_old == v_new
        print("N=%d %-16s cells %.2f ms (%d launches)  windows %.2f ms (%d launches)  x%.2f  %s"
              % (n, name, t_old, k_old, t_new, k_new, t_old / t_new, "same" if same else "MISMATCH"), flush=True)
