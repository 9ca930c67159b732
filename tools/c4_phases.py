"""Phase timings of config 4 (directional, uniform bins) - developer tool, run with SKG_SELECT_DEBUG=1."""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import torch
import skgstat_b200 as skg
from skgstat_b200 import engine as E
from skgstat_b200.data import gaussian_field, random_points
warnings.simplefilter("ignore")
c4 = random_points(100000, 10000.0, 2, seed=4); v4 = gaussian_field(c4, 800.0, seed=4)
mk = lambda: skg.DirectionalVariogram(c4, v4, azimuth=45, tolerance=22.5, bin_func="uniform", n_lags=10, fit_method=None)
mk().experimental
torch.cuda.synchronize()
# wrap engine methods with timers
def wrap(name):
    f = getattr(E.PairEngine, name)
    def g(self, *a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(self, *a, **k)
        torch.cuda.synchronize()
        print("   [call] %-22s %.3f ms" % (name, (time.perf_counter() - t0) * 1e3))
        return r
    setattr(E.PairEngine, name, g)
for nm in ("__init__", "set_values", "ambiguous_pairs", "sq_cell_counts", "hist", "sq_order_stats", "sq_max", "hist_counts"):
    wrap(nm)
t0 = time.perf_counter(); mk().experimental; torch.cuda.synchronize()
print("total %.2f ms" % ((time.perf_counter() - t0) * 1e3))
