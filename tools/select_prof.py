"""One large-N median select for ncu (developer tool)."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import torch
from skgstat_b200.data import random_points
from skgstat_b200.engine import PairEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
c = random_points(n, 10000.0, 2, seed=0)
eng = PairEngine(c, None)
for _ in range(2):
    v, t = eng.sq_order_stats(lambda t: [(t - 1) // 2, t // 2])
torch.cuda.synchronize()
print(v, t)
