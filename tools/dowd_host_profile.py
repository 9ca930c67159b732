"""cProfile of the Dowd sub-sample window select (developer tool)."""
import os, sys, time, warnings, cProfile, pstats
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import torch
import skgstat_b200 as skg
from skgstat_b200.engine import PairEngine
from skgstat_b200._exact import sq_threshold_bits, to_bits
from skgstat_b200.data import gaussian_field, random_points
warnings.simplefilter("ignore")
c = random_points(200000, 10000.0, 2, seed=3); v = gaussian_field(c, 800.0, seed=3)
V = skg.Variogram(c, v, estimator="matheron", n_lags=25, maxlag="median", fit_method=None)
thr = sq_threshold_bits(V.bins)
eng = PairEngine(c, v)
hi = to_bits(eng._dz_bound) + 1
for _ in range(3):
    eng._median_windows(thr, None, hi)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    eng._median_windows(thr, None, hi)
torch.cuda.synchronize()
print("median windows %.3f ms" % ((time.perf_counter() - t0) * 50))
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    eng._median_windows(thr, None, hi)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(25)
st.sort_stats("cumulative").print_stats(30)
