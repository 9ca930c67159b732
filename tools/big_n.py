"""N = 10^6 points (5*10^11 pairs): scale check of the whole variogram path (developer tool)."""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import skgstat_b200 as skg
from skgstat_b200.data import gaussian_field, random_points
import torch
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
c = random_points(n, 30000.0, 2, seed=1); v = gaussian_field(c, 2000.0, seed=1)
t0 = time.perf_counter()
V = skg.Variogram(c, v, n_lags=25, maxlag="median", fit_method=None)
cnt = V.bin_count; exp = V.experimental
torch.cuda.synchronize(); dt = time.perf_counter() - t0
P = n * (n - 1) // 2
print("N=%d pairs=%.3e e2e %.2f s -> %.3e pairs/s; maxlag=%.6f; sum(count)=%d vs P/2=%d (diff %d); exp[0]=%.6g exp[-1]=%.6g"
      % (n, P, dt, P / dt, V.maxlag, cnt.sum(), P // 2, cnt.sum() - P // 2, exp[0], exp[-1]))
assert abs(int(cnt.sum()) - P // 2) <= 2
assert np.all(np.isfinite(exp)) and np.all(cnt > 0)
