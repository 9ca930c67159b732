"""cProfile of the config-2 public call (developer tool): where the host time of the 2 ms goes."""
import os, sys, time, warnings, cProfile, pstats
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import skgstat_b200 as skg
from skgstat_b200.data import gaussian_field, random_points
import torch
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
c = random_points(n, 1000.0, 2, seed=42); v = gaussian_field(c, 150.0, seed=42)
call = lambda: skg.Variogram(c, v, n_lags=25, maxlag="median", fit_method=None).experimental
for _ in range(10):
    call()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    call()
torch.cuda.synchronize()
print("e2e %.3f ms" % ((time.perf_counter() - t0) * 10))
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    call()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
st.sort_stats("cumulative").print_stats(45)
