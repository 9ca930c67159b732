"""Kernel-only timing of skg_pair_hist for several N (developer tool, not the bench)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
from skgstat_b200 import _lib  # noqa: E402
from skgstat_b200._exact import sq_threshold_bits  # noqa: E402
from skgstat_b200.data import gaussian_field, random_points  # noqa: E402
from skgstat_b200.engine import PairEngine, _hptr, _ptr  # noqa: E402

lib = _lib.load()
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [20000, 60000]
stats = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 25
frac = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5   # maxlag as quantile of distances
for n in sizes:
    c = random_points(n, 1000.0, 2, seed=0)
    v = gaussian_field(c, 100.0, seed=0)
    eng = PairEngine(c, v)
    maxlag = {0.5: 521.0, 1.0: 1500.0}.get(frac, 521.0 * frac / 0.5)
    edges = np.linspace(0, maxlag, nl + 1)[1:]
    thr = sq_threshold_bits(edges)
    cnt = torch.zeros(nl, dtype=torch.int64, device="cuda")
    sq = torch.zeros(nl, dtype=torch.float64, device="cuda")
    rt = torch.zeros(nl, dtype=torch.float64, device="cuda")
    scratch = eng._hist_scratch(nl)
    ts = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.skg_pair_hist(_ptr(eng.coords), _ptr(eng.values), 1, 0, eng.n, eng.dim, None, _hptr(thr),
                               nl, stats, None, _ptr(cnt), _ptr(sq), _ptr(rt), _ptr(scratch),
                               scratch.numel(), 0, eng.n_tiles, 1, eng._stream())
        e1.record()
        _lib.check(rc, "hist")
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    P = n * (n - 1) // 2
    inr = int(cnt.sum().item()) / 8 / P
    print("N=%d pairs=%.3e stats=%d nl=%d in-range=%.2f  %.3f ms  %.3e pairs/s  %.2f pairs/clk/SM"
          % (n, P, stats, nl, inr, ms, P / ms * 1e3, P / (ms * 1e-3) / 148 / 1.965e9))
