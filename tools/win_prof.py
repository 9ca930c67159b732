"""One fused pass (window kernel) at N points for ncu (developer tool)."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import torch
import skgstat_b200 as skg
from skgstat_b200 import _lib
from skgstat_b200._exact import sq_threshold_bits
from skgstat_b200.data import gaussian_field, random_points
from skgstat_b200.engine import PairEngine
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
stats = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L = 1000.0 * np.sqrt(n / 20000.0)
c = random_points(n, L, 2, seed=0)
v = gaussian_field(c, L / 10, seed=0)
V = skg.Variogram(c, v, n_lags=25, maxlag="median", fit_method=None)
thr = sq_threshold_bits(V.bins)
eng = PairEngine(c, v)
for _ in range(3):
    out = eng.hist(V.bins, stats, None, thr_bits=thr)
torch.cuda.synchronize()
print(out[0].sum())
