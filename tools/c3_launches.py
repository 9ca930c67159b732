import os, sys, warnings, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import skgstat_b200 as skg
from skgstat_b200.data import gaussian_field, random_points
warnings.simplefilter("ignore")
c = random_points(200000, 10000.0, 2, seed=0); v = gaussian_field(c, 1000.0, seed=0)
f = lambda: skg.Variogram(c, v, estimator="matheron", n_lags=25, maxlag="median", fit_method=None).experimental
f(); f()
