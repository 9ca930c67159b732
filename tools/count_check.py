"""Exact class counts of the count-only fused pass under every path combination (developer tool):
window kernel / first kernel, bulk counting on/off, forced group sizes - all must agree."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import torch
from skgstat_b200.data import random_points
from skgstat_b200.engine import PairEngine
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
L = 10000.0 * np.sqrt(n / 200000.0)
c = random_points(n, L, 2, seed=0)
eng = PairEngine(c, None)
recorded = []
orig = eng.hist_counts


def rec(thr, dirp=None):
    recorded.append(np.array(thr, dtype=np.uint64))
    return orig(thr, dirp)


eng.hist_counts = rec
med = lambda t: [] if t == 0 else [(t - 1) // 2, t // 2]
try:
    eng.sq_order_stats(med)
except Exception as e:
    print("select raised:", e)
eng.hist_counts = orig
combos = [("default", {}), ("bulk off", {"SKG_BULK_OFF": "1"}), ("win G=0", {"SKG_WIN_G": "0"}),
          ("win G=16", {"SKG_WIN_G": "16"}), ("win G=32", {"SKG_WIN_G": "32"}),
          ("win G=16 bulk off", {"SKG_WIN_G": "16", "SKG_BULK_OFF": "1"}),
          ("first kernel", {"SKG_WIN_OFF": "1"}), ("first kernel bulk off", {"SKG_WIN_OFF": "1", "SKG_BULK_OFF": "1"})]
for k, thr in enumerate(recorded):
    print("threshold set %d: %d thresholds" % (k, thr.size))
    ref = None
    for name, env in combos:
        for key in ("SKG_BULK_OFF", "SKG_WIN_G", "SKG_WIN_OFF"):
            os.environ.pop(key, None)
        os.environ.update(env)
        cnt = eng.hist_counts(thr)
        if ref is None:
            ref = cnt
        bad = np.flatnonzero(cnt != ref)
        print("   %-24s total %d  %s" % (name, cnt.sum(), "same" if bad.size == 0 else
                                           "DIFF at classes %s: %s vs %s" % (bad[:6], cnt[bad[:6]], ref[bad[:6]])))
