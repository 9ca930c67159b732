"""Batched sweep (skg_pair_hist_batch) vs one skg_pair_hist per vector (developer tool)."""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import skgstat_b200 as skg
from skgstat_b200 import _lib
from skgstat_b200.data import gaussian_field, random_points
import torch
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
c = random_points(n, 1000.0, 2, seed=0); v = gaussian_field(c, 100.0, seed=0)
V = skg.Variogram(c, v, n_lags=25, maxlag="median", fit_method=None)
rng = np.random.default_rng(0)
draws = np.vstack([v + rng.normal(0, 0.3, n) for _ in range(K)])
eng = V._engine(); edges = V.bins
P = n * (n - 1) // 2
for stat, name in ((_lib.STAT_SUM_SQ, "matheron"), (_lib.STAT_SUM_SQRT, "cressie")):
    eng.hist_batch(edges, draws, stat); torch.cuda.synchronize()
    t0 = time.perf_counter(); eng.hist_batch(edges, draws, stat); torch.cuda.synchronize(); tb = time.perf_counter() - t0
    eng.hist(edges, stat); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for q in range(K):
        eng.set_values(draws[q]); eng.hist(edges, stat)
    torch.cuda.synchronize(); ts = time.perf_counter() - t0
    print("%s N=%d K=%d: batched %.2f ms (%.3e vector-pairs/s) | per-vector %.2f ms (%.3e) | x%.2f"
          % (name, n, K, tb * 1e3, K * P / tb, ts * 1e3, K * P / ts, ts / tb))
# device-only timing of the batch call (inputs resident)
dev = torch.from_numpy(draws).cuda()[:, eng._perm_dev].contiguous()
import ctypes as C
from skgstat_b200.engine import _ptr, _hptr
from skgstat_b200._exact import sq_threshold_bits
thr = sq_threshold_bits(edges); nl = len(edges)
cnt = torch.zeros(nl, dtype=torch.int64, device="cuda"); out = torch.zeros(K * nl, dtype=torch.float64, device="cuda")
scr = torch.empty(int(eng.lib.skg_pair_batch_scratch_bytes(n, nl)), dtype=torch.uint8, device="cuda")
def run():
    return eng.lib.skg_pair_hist_batch(_ptr(eng.coords), _ptr(dev), K, n, 2, None, _hptr(thr), nl, 1, _ptr(cnt), _ptr(out),
                                       _ptr(scr), scr.numel(), 0, eng.n_tiles, 1, eng._stream())
for _ in range(3): run()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("kernel-only batched K=%d: %.3f ms -> %.3e vector-pairs/s (single-vector kernel: ~8e11)" % (K, ms, K * P / (ms * 1e-3)))
