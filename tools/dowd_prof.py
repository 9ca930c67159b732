"""Where the Dowd sweeps spend their time (developer tool): the count+below+histogram sweep with
the real |dz| windows, with empty windows (no histogram atomics), and the fused pass for scale."""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import torch
import skgstat_b200 as skg
from skgstat_b200 import _lib
from skgstat_b200.engine import PairEngine
from skgstat_b200._exact import sq_threshold_bits, to_bits
from skgstat_b200.data import gaussian_field, random_points
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
c = random_points(n, 10000.0, 2, seed=3); v = gaussian_field(c, 800.0, seed=3)
V = skg.Variogram(c, v, estimator="matheron", n_lags=25, maxlag="median", fit_method=None)
edges = V.bins
thr = sq_threshold_bits(edges)
eng = PairEngine(c, v)
hi = to_bits(eng._dz_bound) + 1
w_lo0, w_hi0 = eng._median_windows(thr, None, hi)
w_lo, w_hi = eng._merge_dz_windows(w_lo0, w_hi0, hi)
print('distinct windows: %d of %d classes' % (len(set(zip(w_lo.tolist(), w_hi.tolist()))), len(w_lo)))
def timed(label, fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    print("%-46s %.2f ms" % (label, (time.perf_counter() - t0) / reps * 1e3), flush=True)
    return r
timed("fused pass cressie", lambda: eng.hist(edges, _lib.STAT_SUM_SQRT, None, thr_bits=thr))
timed("fused pass matheron", lambda: eng.hist(edges, _lib.STAT_SUM_SQ, None, thr_bits=thr))
cnt, below, hist, sc = timed("dz windows sweep, real windows (4096 buckets)", lambda: eng.dz_windows(thr, w_lo, w_hi, 4096, None))
print("   in-window share %.3f" % (hist.sum() / cnt.sum()))
timed("dz windows sweep, unmerged windows (4096 buckets)", lambda: eng.dz_windows(thr, w_lo0, w_hi0, 4096, None))
timed("dz windows sweep, empty windows", lambda: eng.dz_windows(thr, w_lo, w_lo, 4096, None))
timed("hist_below (count + below, no windows)", lambda: eng.hist_below(thr, w_lo.view(np.float64), None))
one_lo = np.full_like(w_lo, w_lo[6:].min()); one_hi = np.full_like(w_hi, w_hi[6:].max())
timed("dz windows sweep, ONE window for all classes", lambda: eng.dz_windows(thr, one_lo, one_hi, 4096, None))
timed("dz windows sweep, ONE empty window for all", lambda: eng.dz_windows(thr, one_lo, one_lo, 4096, None))
PairEngine.WINDOW_MERGE_GROWTH = 2.0
m_lo, m_hi = eng._merge_dz_windows(w_lo0, w_hi0, hi)
print('growth 2.0: distinct windows: %d' % len(set(zip(m_lo.tolist(), m_hi.tolist()))))
c2_, b2_, h2_, _ = timed("dz windows sweep, growth 2.0", lambda: eng.dz_windows(thr, m_lo, m_hi, 4096, None))
print("   in-window share %.3f" % (h2_.sum() / c2_.sum()))
print("class  lo0      hi0      | merged lo  hi")
for g in range(len(w_lo)):
    print("%3d  %.5f %.5f | %.5f %.5f" % (g, w_lo0.view(np.float64)[g], w_hi0.view(np.float64)[g], w_lo.view(np.float64)[g], w_hi.view(np.float64)[g]))
