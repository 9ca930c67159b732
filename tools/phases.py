"""Per-call timings of the public variogram call - developer tool.
   python tools/phases.py c2|c3|c3c|c3d|c4   (optionally SKG_SELECT_DEBUG=1 for the sweeps inside)"""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import torch
import skgstat_b200 as skg
from skgstat_b200 import engine as E
from skgstat_b200.data import gaussian_field, random_points
warnings.simplefilter("ignore")
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c2":
    c = random_points(20000, 1000.0, 2, seed=42); v = gaussian_field(c, 150.0, seed=42)
    mk = lambda: skg.Variogram(c, v, n_lags=25, maxlag="median", fit_method=None)
elif which in ("c3", "c3c", "c3d"):
    c = random_points(200000, 10000.0, 2, seed=3); v = gaussian_field(c, 800.0, seed=3)
    est = {"c3": "matheron", "c3c": "cressie", "c3d": "dowd"}[which]
    mk = lambda: skg.Variogram(c, v, estimator=est, n_lags=25, maxlag="median", fit_method=None)
else:
    c = random_points(100000, 10000.0, 2, seed=4); v = gaussian_field(c, 800.0, seed=4)
    mk = lambda: skg.DirectionalVariogram(c, v, azimuth=45, tolerance=22.5, bin_func="uniform", n_lags=10, fit_method=None)
for _ in range(3):
    mk().experimental
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    mk().experimental
torch.cuda.synchronize()
print("%s e2e %.3f ms (no instrumentation)" % (which, (time.perf_counter() - t0) * 200))
depth = [0]
def wrap(name):
    f = getattr(E.PairEngine, name)
    def g(self, *a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); depth[0] += 1
        r = f(self, *a, **k)
        torch.cuda.synchronize(); depth[0] -= 1
        print("   %s[call] %-22s %.3f ms  (n=%d)" % ("  " * depth[0], name, (time.perf_counter() - t0) * 1e3, self.n))
        return r
    setattr(E.PairEngine, name, g)
for nm in ("__init__", "set_values", "ambiguous_pairs", "sq_cell_counts", "hist", "sq_order_stats", "sq_max",
           "hist_counts", "dz_medians", "dz_windows", "_window_cell_counts", "_median_windows"):
    if hasattr(E.PairEngine, nm):
        wrap(nm)
t0 = time.perf_counter(); mk().experimental; torch.cuda.synchronize()
print("total %.2f ms (instrumented)" % ((time.perf_counter() - t0) * 1e3))
