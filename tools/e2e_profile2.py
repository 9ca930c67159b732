import cProfile, os, pstats, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import skgstat_b200 as skg
from skgstat_b200.data import gaussian_field, random_points
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
c = random_points(n, 1000.0, 2, seed=0); v = gaussian_field(c, 100.0, seed=0)
kw = dict(n_lags=25, maxlag="median", fit_method=None)
for _ in range(3): skg.Variogram(c, v, **kw).experimental
pr = cProfile.Profile(); pr.enable()
for _ in range(10): skg.Variogram(c.copy(), v.copy(), **kw).experimental
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
