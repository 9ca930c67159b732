"""One batched sweep (K2b) at the bench config, for ncu captures (developer tool)."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import skgstat_b200 as skg
from skgstat_b200.data import gaussian_field, random_points
warnings.simplefilter("ignore")
c = random_points(20000, 1000.0, 2, seed=0); v = gaussian_field(c, 100.0, seed=0)
V = skg.Variogram(c, v, n_lags=25, maxlag="median", fit_method=None)
rng = np.random.default_rng(0)
draws = np.vstack([v + rng.normal(0, 0.3, v.size) for _ in range(8)])
V.experimental_batch(draws); V.experimental_batch(draws)
