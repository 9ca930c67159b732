"""cProfile of the end-to-end class call (developer tool)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import skgstat_b200 as skg  # noqa: E402
from skgstat_b200.data import gaussian_field, random_points  # noqa: E402
import torch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
c = random_points(n, 1000.0, 2, seed=0)
v = gaussian_field(c, 100.0, seed=0)
kw = dict(n_lags=25, maxlag="median", fit_method=None)
for _ in range(3):
    skg.Variogram(c, v, **kw).experimental
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    skg.Variogram(c.copy(), v.copy(), **kw).experimental
torch.cuda.synchronize()
print("e2e ms/call:", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    skg.Variogram(c.copy(), v.copy(), **kw).experimental
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
