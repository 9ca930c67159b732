"""Fused-pass kernel timings with the register-window path on/off and forced group sizes
(developer tool; CUDA events around engine.hist's C-ABI call)."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import torch
import skgstat_b200 as skg
from skgstat_b200 import _lib
from skgstat_b200._exact import sq_threshold_bits
from skgstat_b200.data import gaussian_field, random_points
from skgstat_b200.engine import PairEngine, _hptr, _ptr
warnings.simplefilter("ignore")
lib = _lib.load()


def kernel_ms(eng, thr, stats, reps=5):
    nl = int(thr.size)
    cnt = torch.zeros(nl, dtype=torch.int64, device="cuda")
    a = torch.zeros(nl, dtype=torch.float64, device="cuda")
    b = torch.zeros(nl, dtype=torch.float64, device="cuda")
    scratch = eng._hist_scratch(nl)
    ts = []
    for i in range(reps + 2):
        cnt.zero_(); a.zero_(); b.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.skg_pair_hist(_ptr(eng.coords), _ptr(eng.values), 1, 0, eng.n, eng.dim, None, _hptr(thr), nl,
                               stats, None, _ptr(cnt), _ptr(a), _ptr(b), _ptr(scratch), scratch.numel(),
                               eng.t0, eng.t1, eng.tstride, eng._stream())
        e1.record()
        _lib.check(rc, "skg_pair_hist")
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    ctr = scratch[:128].view(torch.int64).cpu().numpy()
    g = int(ctr[4])
    dd = ctr[5:8].view(np.float64)
    print("      plan: diag16 %.3f diag32 %.3f width %.3f -> ratios %.2f %.2f" % (dd[0], dd[1], dd[2], dd[0] / dd[2], dd[1] / dd[2]))
    if ctr[8:14].any():   # debug build: warp-groups per path (dead, W=2, W=3, per-pair, diagonal items, culled items)
        print("      paths: dead %d  W2 %d  W3 %d  per-pair %d  diag items %d  culled warp-items %d  (list %d of %d)"
              % (ctr[8], ctr[9], ctr[10], ctr[11], ctr[12], ctr[13], ctr[0], ctr[3]))
    return min(ts), cnt.cpu().numpy(), a.cpu().numpy(), b.cpu().numpy(), g


def main():
    sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["20000", "60000", "200000"])]
    for n in sizes:
        L = 1000.0 * np.sqrt(n / 20000.0)
        c = random_points(n, L, 2, seed=0)
        v = gaussian_field(c, L / 10, seed=0)
        V = skg.Variogram(c, v, n_lags=25, maxlag="median", fit_method=None)
        thr = sq_threshold_bits(V.bins)
        eng = PairEngine(c, v)
        P = n * (n - 1) // 2
        for stats, name in ((1, "matheron"), (2, "cressie")):
            os.environ["SKG_WIN_OFF"] = "1"
            t0, c0, a0, b0, _ = kernel_ms(eng, thr, stats)
            del os.environ["SKG_WIN_OFF"]
            print("N=%d %-8s old kernel      %.3f ms  %.3e pairs/s" % (n, name, t0, P / t0 * 1e3), flush=True)
            for g in ("auto", "0", "8", "16", "32"):
                if g == "auto":
                    os.environ.pop("SKG_WIN_G", None)
                else:
                    os.environ["SKG_WIN_G"] = g
                t1, c1, a1, b1, gg = kernel_ms(eng, thr, stats)
                ok = np.array_equal(c0, c1) and np.allclose(a0, a1, rtol=1e-13) and np.allclose(b0, b1, rtol=1e-13)
                print("N=%d %-8s win G=%-4s (%2d) %.3f ms  %.3e pairs/s  x%.2f  %s"
                      % (n, name, g, gg, t1, P / t1 * 1e3, t0 / t1, "same" if ok else "MISMATCH"), flush=True)
            os.environ.pop("SKG_WIN_G", None)


if __name__ == "__main__":
    main()
