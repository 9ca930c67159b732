import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200")); sys.path.insert(0, ROOT)
import skgstat_b200 as skg
from oracle.fields import gaussian_field, random_points
warnings.simplefilter("ignore")
c = random_points(400, 100.0, 2, seed=1); v = gaussian_field(c, 12.0, seed=1)
V = skg.Variogram(c, v, n_lags=10, maxlag="median", model="exponential")
rng = np.random.default_rng(1234)
draws = np.vstack([rng.normal(v, 0.3, size=400) for _ in range(40)])
print("draw std across iterations", draws.std(axis=0)[:3])
E = V.experimental_batch(draws)
print(E[:3, :3]); print(E[37:, :3])
args = {k: val for k, val in V.describe()["params"].items() if k not in ("values", "verbose")}
print(args)
T = skg.Variogram(coordinates=V._X, values=draws[0], **args)
print("template bins", T.bins[:3], "V bins", V.bins[:3])
E2 = T.experimental_batch(draws)
print(E2[:3, :3])
