"""Kernel-only timing of skg_krige (developer tool)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
from skgstat_b200.data import gaussian_field, random_points  # noqa: E402
from skgstat_b200.engine import KrigeEngine, to_device  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
maxp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
radius = float(sys.argv[3]) if len(sys.argv) > 3 else 464.0
n = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
c = random_points(n, 1000.0, 2, seed=5)
v = gaussian_field(c, 100.0, seed=5)
eng = KrigeEngine(c, v, radius)
gx, gy = np.mgrid[0:1000:complex(0, side), 0:1000:complex(0, side)].reshape(2, -1)
m = gx.size
tg = to_device(np.vstack((gx, gy)), eng.device)
mp = np.array([radius, 4.0, 0.0, 0.0])
ts = []
for i in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    z, s, st = eng.transform_device(tg, m, "exponential", mp, min(5, maxp), maxp)
    e1.record()
    torch.cuda.synchronize()
    if i:
        ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
print("targets=%d samples=%d max_points=%d radius=%g grid h=%.1f: %.2f ms  %.3e points/s  nan=%d"
      % (m, n, maxp, radius, eng.grid[3], ms, m / ms * 1e3, int(torch.isnan(z).sum())))
