"""e2e timings of the BASELINE.json configurations through the class API (developer tool)."""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import skgstat_b200 as skg
from skgstat_b200.data import gaussian_field, random_points
import torch
warnings.simplefilter("ignore")
def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
    return min(t)
c = random_points(200000, 10000.0, 2, seed=0); v = gaussian_field(c, 1000.0, seed=0)
P = 200000 * 199999 // 2
for est in ("matheron", "cressie", "dowd"):
    dt = timeit(lambda: skg.Variogram(c, v, estimator=est, n_lags=25, maxlag="median", fit_method=None).experimental)
    print("C3 N=200k %-9s e2e %.1f ms  %.3e pairs/s" % (est, dt * 1e3, P / dt))
c4 = random_points(100000, 10000.0, 2, seed=4); v4 = gaussian_field(c4, 800.0, seed=4)
P4 = 100000 * 99999 // 2
dt = timeit(lambda: skg.DirectionalVariogram(c4, v4, azimuth=45, tolerance=22.5, bin_func="uniform", n_lags=10, fit_method=None).experimental)
print("C4 N=100k directional uniform e2e %.1f ms  %.3e pairs/s" % (dt * 1e3, P4 / dt))
c2 = random_points(20000, 1000.0, 2, seed=0); v2 = gaussian_field(c2, 100.0, seed=0)
dt = timeit(lambda: skg.Variogram(c2, v2, n_lags=25, maxlag="median", fit_method=None).experimental, 5)
print("C2 N=20k matheron e2e %.2f ms" % (dt * 1e3))
