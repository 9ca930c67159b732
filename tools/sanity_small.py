"""Small end-to-end pass over every kernel family (for compute-sanitizer)."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
sys.path.insert(0, ROOT)
import skgstat_b200 as skg
from skgstat_b200.data import gaussian_field, random_points
warnings.simplefilter("ignore")
c = random_points(700, 100.0, 2, seed=1); v = gaussian_field(c, 10.0, seed=1)
for est in ("matheron", "cressie", "dowd"):
    V = skg.Variogram(c, v, estimator=est, n_lags=12, maxlag="median", bin_func="uniform")
    print(est, V.bin_count.sum(), float(V.experimental[3]))
V = skg.Variogram(c, v, n_lags=120, fit_method=None); print("many lags", V.bin_count.sum())
DV = skg.DirectionalVariogram(c, v, azimuth=30, tolerance=40, bin_func="uniform", fit_method=None)
print("dir", DV.bin_count.sum(), int(DV._direction_mask().sum()))
c3 = random_points(300, 50.0, 3, seed=2); v3 = gaussian_field(c3, 8.0, seed=2)
V3 = skg.Variogram(c3, v3, n_lags=8, fit_method=None); print("3d", V3.bin_count.sum(), V3.distance.shape)
Vk = skg.Variogram(c, v, model="exponential", n_lags=15, maxlag="median")
ok = skg.OrdinaryKriging(Vk, min_points=3, max_points=33)
gx, gy = np.mgrid[-10:110:25j, -10:110:25j].reshape(2, -1)
z = ok.transform(gx, gy); print("krige", np.isnan(z).sum(), float(np.nanmean(z)))
print("cv", Vk.cross_validate(n=100, seed=1))
print("SANITY DONE")
