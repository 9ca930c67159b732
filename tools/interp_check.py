import os, sys, numpy as np
sys.path.insert(0, "/root/repo/scikit-gstat_b200")
from skgstat_b200.data import random_points, gaussian_field
from skgstat_b200.engine import PairEngine
from skgstat_b200 import binning
for n, L, seed in ((200000, 10000.0, 0), (100000, 10000.0, 4), (120000, 5000.0, 9), (300000, 20000.0, 3)):
    c = random_points(n, L, 2, seed=seed)
    if seed == 9:   # clustered points: a rougher distance distribution
        rng = np.random.default_rng(1); c = c * 0.2 + rng.integers(0, 5, size=(n, 1)) * (L / 5.0)
    eng = PairEngine(c, None)
    for name, fn in (("median", lambda t: [(t - 1) // 2, t // 2]), ("q33", lambda t: list(binning.percentile_ranks(t, 33)[:2])), ("q90", lambda t: list(binning.percentile_ranks(t, 90)[:2]))):
        print(n, name, flush=True)
        eng.sq_order_stats(fn)
