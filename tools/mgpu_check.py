"""torchrun check (N GPUs): the sharded path equals the single-GPU path bit for bit on
counts/edges and to rounding on sums; kriging slices concatenate to the full field."""
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import skgstat_b200 as skg  # noqa: E402
from skgstat_b200.data import gaussian_field, random_points  # noqa: E402

rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
warnings.simplefilter("ignore")
c = random_points(7000, 1000.0, 2, seed=3)
v = gaussian_field(c, 90.0, seed=3)
ok_all = True
for est, bf in (("matheron", "even"), ("cressie", "uniform"), ("dowd", "even")):
    A = skg.Variogram(c, v, estimator=est, bin_func=bf, n_lags=20, maxlag="median", fit_method=None)
    B = skg.Variogram(c, v, estimator=est, bin_func=bf, n_lags=20, maxlag="median", fit_method=None,
                      process_group=dist.group.WORLD)
    good = (np.array_equal(A.bins, B.bins) and np.array_equal(A.bin_count, B.bin_count)
            and np.allclose(A.experimental, B.experimental, rtol=1e-13, atol=0))
    ok_all &= good
    if rank == 0:
        print(est, bf, "sharded == single:", good)
A = skg.DirectionalVariogram(c, v, azimuth=45, tolerance=22.5, bin_func="uniform", fit_method=None)
B = skg.DirectionalVariogram(c, v, azimuth=45, tolerance=22.5, bin_func="uniform", fit_method=None,
                             process_group=dist.group.WORLD)
good = np.array_equal(A.bins, B.bins) and np.array_equal(A.bin_count, B.bin_count)
ok_all &= good
if rank == 0:
    print("directional sharded == single:", good)
# rows added after the first scaling run: multi-column values, batched sweep, histogram-rule
# binners, extra estimators - all through the sharded tile ranges + one all-reduce
G = dist.group.WORLD
v2 = gaussian_field(c, 40.0, seed=9)
rng = np.random.default_rng(5)
draws = np.vstack([v + rng.normal(0, 0.2, v.size) for _ in range(6)])
checks = []
for mode, vals in (("cross", np.column_stack((v, v2))), ("aggregate", np.column_stack((v, v2, v * 0.5)))):
    for est in ("matheron", "dowd"):
        A = skg.Variogram(c, vals, estimator=est, multivariate=mode, n_lags=12, maxlag="median", fit_method=None)
        B = skg.Variogram(c, vals, estimator=est, multivariate=mode, n_lags=12, maxlag="median", fit_method=None,
                          process_group=G)
        checks.append(("%s/%s" % (mode, est), np.array_equal(A.bin_count, B.bin_count)
                       and np.allclose(A.experimental, B.experimental, rtol=1e-13, atol=0)))
A = skg.Variogram(c, v, n_lags=20, maxlag="median", fit_method=None)
B = skg.Variogram(c, v, n_lags=20, maxlag="median", fit_method=None, process_group=G)
checks.append(("batched sweep", np.allclose(A.experimental_batch(draws), B.experimental_batch(draws), rtol=1e-13, atol=0)))
for bf in ("sturges", "doane", "scott"):
    A = skg.Variogram(c[:2500], v[:2500], bin_func=bf, maxlag="median", fit_method=None)
    B = skg.Variogram(c[:2500], v[:2500], bin_func=bf, maxlag="median", fit_method=None, process_group=G)
    checks.append(("bin_func=" + bf, np.array_equal(A.bins, B.bins) and np.array_equal(A.bin_count, B.bin_count)))
for est in ("minmax", "percentile", "entropy"):
    A = skg.Variogram(c, v, estimator=est, n_lags=10, maxlag="median", fit_method=None)
    B = skg.Variogram(c, v, estimator=est, n_lags=10, maxlag="median", fit_method=None, process_group=G)
    checks.append(("estimator=" + est, np.allclose(A.experimental, B.experimental, rtol=1e-12, atol=0)))
for name, good in checks:
    ok_all &= bool(good)
    if rank == 0:
        print(name, "sharded == single:", bool(good))
V = skg.Variogram(c, v, model="exponential", n_lags=20, maxlag="median")
gx, gy = np.mgrid[0:1000:150j, 0:1000:150j].reshape(2, -1)
k1 = skg.OrdinaryKriging(V, min_points=5, max_points=32)
k2 = skg.OrdinaryKriging(V, min_points=5, max_points=32, process_group=dist.group.WORLD)
z1, z2 = k1.transform(gx, gy), k2.transform(gx, gy)
good = np.array_equal(z1, z2, equal_nan=True) and np.array_equal(k1.sigma, k2.sigma, equal_nan=True)
ok_all &= good
if rank == 0:
    print("kriging sharded == single:", good)
    print("MGPU_CHECK", "PASS" if ok_all else "FAIL")
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
