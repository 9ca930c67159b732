#!/usr/bin/env python
"""bench.py - variogram pairs/s (fused distance + lag bin + estimator) and kriged points/s on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...        (one rank per GPU, NCCL)

Own arm.  HEADLINE = BASELINE.json config 3, the configuration its metric is quoted on for 1/2/4/8
GPUs: 2-D synthetic Gaussian field, N = 200 000 points (19 999 900 000 pairs), estimator cressie,
n_lags = 25, maxlag = median.  The SAME N at every world size (strong scaling): the tile triangle is
dealt round-robin to the ranks, the per-lag partials meet in ONE all-reduce per step.
  step   = one fused pass (skg_pair_hist: distance -> lag class -> count and sum sqrt|dz| of every
           pair), inputs resident in HBM; `value` = pairs / max-over-ranks step time
  e2e    = Variogram(coords, values, 'cressie', n_lags=25, maxlag='median').experimental through the
           class API with HOST buffers: upload, Z-order sort, exact median select, edges, fused pass,
           download - every step
Extra keys (each with its own kernel ms / e2e ms / roofline):
  config2       N = 20 000, matheron (the configuration a single B200 is quoted on)
  config3_dowd  N = 200 000, dowd (per-lag exact median select instead of a sum)
  config4       DirectionalVariogram N = 100 000, azimuth 45, tolerance 22.5, uniform bins
  kriging       config 5: OrdinaryKriging.transform, 2000 x 2000 grid from 10 000 samples, 5..64
  batched       32 value vectors over config 2's coordinates in one batched sweep (N = 1 only)
  cpu_baseline  the stock reference (oracle/_ref or /root/reference, unmodified, through its own
                Variogram / OrdinaryKriging classes) on a bounded sample, host cores (N = 1 only)
Roofline: FP64 ALU (neither path is HBM- or tensor-bound); the denominator is a DFMA
chain measured in this run (MEASURED_PEAKS.json holds no FP64 figure).

Reference arm (--impl reference): the UNMODIFIED reference's own constructor on a bounded sample
of the headline workload, on every host core (independent replicas: the reference is
single-threaded by construction).  Falls back to the oracle port only if no reference copy exists.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "scikit-gstat_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "variogram pairs/s (fused dist+bin+estimator)"
UNIT = "pairs/s"
N_LAGS = 25
# FP64 operations per pair (2-D): 3 sub + 3 mul + 2 add + 1 sqrt-equivalent for
# matheron; cressie adds sqrt|dz| and its accumulate; the directional mask ~6 mul/add
FLOPS = {"matheron": 9.0, "cressie": 11.0, "directional": 15.0, "dowd": 8.0}
C3 = dict(n=200000, extent=10000.0, corr=1000.0, seed=0)
C2 = dict(n=20000, extent=1000.0, corr=100.0, seed=0)
C4 = dict(n=100000, extent=10000.0, corr=800.0, seed=4)
HEADLINE = ("2-D synthetic Gaussian field N=200000 (19999900000 pairs), cressie, n_lags=25, "
            "maxlag=median (BASELINE.json config 3)")
# from the committed ncu --set full captures of the dominant kernels (profiles/)
NCU = {
    "cressie": {"fp64_pipe_pct": 63.9, "issue_active_pct": 59.8, "smem_wavefronts_pct": 21.4,
                "traffic_bytes": 6699264,
                "from": "profiles/r2_pair_hist_win_cressie_ncu_summary.txt (N=200k, 15.80 ms under ncu)"},
    "matheron": {"fp64_pipe_pct": 58.7, "issue_active_pct": 61.4, "smem_wavefronts_pct": 30.7,
                 "traffic_bytes": 6692352,
                 "from": "profiles/r2_pair_hist_win_ncu_summary.txt (N=200k, 11.07 ms under ncu)"},
    "kriging": {"fp64_pipe_pct": 7.3, "issue_active_pct": 40.6, "smem_wavefronts_pct": 40.6,
                "traffic_bytes": 244.2e6,
                "from": "profiles/r2_krige_chol_ncu_summary.txt (400x400 targets, 8.66 ms under ncu)"},
}


def make_inputs(n, extent, corr, seed):
    from skgstat_b200.data import gaussian_field, random_points
    c = random_points(n, extent, 2, seed=seed)
    return c, gaussian_field(c, corr, seed=seed)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed regions run."""

    def __init__(self, index, period=float(os.environ.get("SKG_CLOCK_PERIOD", "0.02"))):
        super().__init__(daemon=True)
        self.period = period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------ reference / CPU legs

def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _ref_variogram_once(args):
    """One stock-reference constructor + experimental (runs in a worker process)."""
    n, estimator, seed = args
    sys.path.insert(0, ROOT)
    from oracle import ref_harness
    skgstat = ref_harness.import_reference()
    from oracle.fields import gaussian_field, random_points
    c = random_points(n, C3["extent"] * np.sqrt(n / C3["n"]), 2, seed=seed)
    v = gaussian_field(c, C3["corr"] * np.sqrt(n / C3["n"]), seed=seed)
    t0 = time.perf_counter()
    V = skgstat.Variogram(c, v, estimator=estimator, n_lags=N_LAGS, maxlag="median")
    e = V.experimental
    dt = time.perf_counter() - t0
    assert np.all(np.isfinite(e))
    return dt


def reference_variogram_rate(n_sample, estimator="cressie", replicas=1, repeats=1):
    """pairs/s of the UNMODIFIED reference's own Variogram(...) constructor (double preprocessing and
    fit included, Variogram.py:31-405) on n_sample points; `replicas` independent processes at a
    time (the reference is single-threaded).  Returns (aggregate pairs/s, seconds per call, kind)."""
    from oracle import ref_harness
    pairs = n_sample * (n_sample - 1) // 2
    if not ref_harness.available():      # no reference copy on this box: the oracle port, and say so
        from oracle import variogram as ov
        from oracle.fields import gaussian_field, random_points
        c = random_points(n_sample, 1000.0, 2, seed=1)
        v = gaussian_field(c, 100.0, seed=1)
        t0 = time.perf_counter()
        ov.dense_variogram(c, v, estimator, "even", N_LAGS, "median")
        dt = time.perf_counter() - t0
        return pairs / dt, dt, "port", 1
    if replicas <= 1:
        dts = [_ref_variogram_once((n_sample, estimator, 1)) for _ in range(repeats)]
        return pairs / min(dts), min(dts), "reference", 1
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(replicas) as pool:
        t0 = time.perf_counter()
        dts = pool.map(_ref_variogram_once, [(n_sample, estimator, 1 + r) for r in range(replicas)])
        wall = time.perf_counter() - t0
    # aggregate throughput of the replicas over the slowest one's compute time (imports excluded)
    return replicas * pairs / max(dts), max(dts), "reference", replicas


def reference_kriging_rate(m_sample=1500):
    """points/s of the reference's OrdinaryKriging.transform (n_jobs=1, Kriging.py:405-474) on a random
    subset of config 5's grid nodes."""
    from oracle import ref_harness
    from oracle.fields import gaussian_field, random_points
    c = random_points(10000, 1000.0, 2, seed=5)
    v = gaussian_field(c, 100.0, seed=5)
    gx, gy = np.mgrid[0:1000:2000j, 0:1000:2000j].reshape(2, -1)
    sub = np.random.default_rng(0).choice(gx.size, m_sample, replace=False)
    if not ref_harness.available():
        return None
    skgstat = ref_harness.import_reference()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        V = skgstat.Variogram(c, v, model="exponential", n_lags=25, maxlag="median")
        ok = skgstat.OrdinaryKriging(V, min_points=5, max_points=64, mode="exact")
        t0 = time.perf_counter()
        z = ok.transform(gx[sub], gy[sub])
        dt = time.perf_counter() - t0
    return {"value": m_sample / dt, "unit": "points/s", "cores": 1, "kind": "reference",
            "sample": "reference OrdinaryKriging.transform on %d random nodes of the 2000x2000 grid, %.1f s"
                      % (m_sample, dt)}, sub, z


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    replicas = max(1, min(cores, 32))
    # size the sample so that warm-up + steps stay within a few minutes: t ~ N^2
    probe_n = 2500
    _, probe_dt, kind, _ = reference_variogram_rate(probe_n, "cressie")
    budget = 150.0 / max(1, args.warmup + args.steps)
    n_sample = int(min(8000, max(1500, probe_n * np.sqrt(max(budget, 0.2) / max(probe_dt, 1e-3)))))
    times, rates = [], []
    for i in range(args.warmup + args.steps):
        rate, dt, kind, used = reference_variogram_rate(n_sample, "cressie", replicas=replicas)
        if i >= args.warmup:
            times.append(dt)
            rates.append(rate)
    ms = 1e3 * float(np.mean(times))
    value = float(np.mean(rates))
    pairs = n_sample * (n_sample - 1) // 2
    sample = ("%s Variogram(coords, values, 'cressie', n_lags=25, maxlag='median') + .experimental on "
              "N=%d points (%d pairs) of the N=200000 workload per step, %d independent replicas on %d "
              "host cores (%s)" % ("stock reference" if kind == "reference" else "oracle port (no reference copy on this box)",
                                   n_sample, pairs, used, cores, cpu_model()))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": HEADLINE, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample,
                         "cpu_model": cpu_model(), "cpu_count": cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- own arm

class Ctx:
    pass


def fused_leg(X, c, v, estimator, steps, warmup, edges=None, dirp=None):
    """Device-timed fused pass over resident inputs: (step ms through engine.hist incl. the
    all-reduce and the result download, kernel ms of the bare C-ABI call, sweep info, launches)."""
    import torch
    from skgstat_b200 import _lib
    from skgstat_b200._exact import sq_threshold_bits
    from skgstat_b200.engine import PairEngine, _hptr, _ptr
    eng = PairEngine(c, v, group=X.group)
    thr = sq_threshold_bits(edges)
    stat = _lib.STAT_SUM_SQRT if estimator == "cressie" else _lib.STAT_SUM_SQ
    for _ in range(max(warmup, 3)):
        eng.hist(edges, stat, dirp, thr_bits=thr)
    X.barrier()
    l0 = eng.launches
    evs = []
    for _ in range(steps):
        X.flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cnt, ssq, srt = eng.hist(edges, stat, dirp, thr_bits=thr)
        e1.record()
        evs.append((e0, e1))
    X.barrier()
    launches = eng.launches - l0
    step_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    assert int(cnt.sum()) > 0
    nl = len(edges)
    cnt_d = torch.zeros(nl, dtype=torch.int64, device="cuda")
    a_d = torch.zeros(nl, dtype=torch.float64, device="cuda")
    b_d = torch.zeros(nl, dtype=torch.float64, device="cuda")
    scratch = eng._hist_scratch(nl)
    kev = []
    for i in range(3 + steps):
        X.flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = X.lib.skg_pair_hist(_ptr(eng.coords), _ptr(eng.values), 1, 0, eng.n, eng.dim, _hptr(dirp), _hptr(thr),
                                 nl, stat, None, _ptr(cnt_d), _ptr(a_d), _ptr(b_d), _ptr(scratch),
                                 scratch.numel(), eng.t0, eng.t1, eng.tstride, eng._stream())
        e1.record()
        _lib.check(rc, "skg_pair_hist")
        if i >= 3:
            kev.append((e0, e1))
    X.barrier()
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    kept, cand, slots = eng.last_sweep_info()
    ctr = scratch[:64].view(torch.int64).cpu().numpy()
    return step_ms, kern_ms, (kept, cand, slots, int(ctr[4])), launches, (cnt, ssq, srt)


def dowd_sweep_leg(X, c, v, edges, steps):
    """Device time of the dominant kernel of the Dowd path: skg_pair_dz_windows (per-lag pair count,
    count below the lag's median window, in-window |dz| histogram) with the windows the engine
    derives from its sub-sample.  Returns (kernel ms, in-window share of the in-range pairs)."""
    import torch
    from skgstat_b200._exact import sq_threshold_bits, to_bits
    from skgstat_b200.engine import PairEngine, _hptr, _ptr
    from skgstat_b200.select import _scale_for
    eng = PairEngine(c, v, group=X.group)
    thr = sq_threshold_bits(edges)
    nl, nbk = len(edges), 1 << 12
    hi = to_bits(eng._dz_bound) + 1
    w_lo, w_hi = eng._merge_dz_windows(*eng._median_windows(thr, None, hi), hi)
    sc = np.array([_scale_for(int(a), int(b), nbk) if b > a else 0.0 for a, b in zip(w_lo, w_hi)])
    cnt = torch.zeros(nl, dtype=torch.int64, device="cuda")
    below = torch.zeros(nl, dtype=torch.float64, device="cuda")
    hist = torch.zeros(nl * nbk, dtype=torch.int64, device="cuda")
    scratch = eng._hist_scratch(nl)
    evs = []
    for i in range(3 + steps):
        X.flush.zero_()
        cnt.zero_(); below.zero_(); hist.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = X.lib.skg_pair_dz_windows(_ptr(eng.coords), _ptr(eng.values), 1, 0, eng.n, eng.dim, None, _hptr(thr),
                                       nl, _hptr(w_lo), _hptr(w_hi), _hptr(sc), nbk, _ptr(cnt), _ptr(below),
                                       _ptr(hist), _ptr(scratch), scratch.numel(), eng.t0, eng.t1, eng.tstride,
                                       eng._stream())
        e1.record()
        assert rc == 0
        if i >= 3:
            evs.append((e0, e1))
    X.barrier()
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    inside = float(hist.sum().item()) / max(float(cnt.sum().item()), 1.0)
    return ms, inside, float(cnt.sum().item())


def e2e_leg(X, make, steps, warm=3):
    """Wall time of the public call with host buffers (H2D + D2H inside), max over ranks."""
    import torch
    from skgstat_b200.engine import Traffic
    for _ in range(warm):
        out = make()
    X.barrier()
    Traffic.reset()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = make()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return X.max_over_ranks(dt), Traffic.h2d // steps, Traffic.d2h // steps, out


def roofline(X, flops_per_pair, pairs_rank, info, kern_ms, kernel, ncu=None):
    kept, cand, slots, group = info
    evaluated = min(float(kept) * slots, pairs_rank) if cand else pairs_rank
    achieved = flops_per_pair * evaluated / (kern_ms * 1e-3) / 1e12
    out = {"bound": "fp64", "achieved": achieved, "peak": X.peak_tflops, "unit": "TFLOP/s",
           "frac": achieved / X.peak_tflops, "traffic": None, "kernel": kernel, "kernel_ms": kern_ms,
           "pairs_evaluated": evaluated, "pairs_total": pairs_rank,
           "culled_fraction": 1.0 - evaluated / pairs_rank,
           "achieved_counting_culled_pairs": flops_per_pair * pairs_rank / (kern_ms * 1e-3) / 1e12,
           "flops_per_pair": flops_per_pair, "window_group_points": group,
           "peak_source": "DFMA chain measured in this run (skg_fp64_peak); MEASURED_PEAKS.json has no "
                          "FP64 figure"}
    if ncu:
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this
        # configuration, from the committed capture named in ncu_from
        out["traffic"] = ncu["traffic_bytes"]
        out["fp64_pipe_pct"] = ncu["fp64_pipe_pct"]
        out["issue_active_pct"] = ncu["issue_active_pct"]
        out["smem_wavefronts_pct"] = ncu["smem_wavefronts_pct"]
        out["ncu_from"] = ncu["from"]
    return out


def run_own(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    X = Ctx()
    X.group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import skgstat_b200 as skg
    from skgstat_b200 import _lib

    X.lib = _lib.load()
    X.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    X.barrier, X.max_over_ranks = barrier, max_over_ranks
    pk_ms, pk_fma = C.c_double(), C.c_double()
    X.peak_tflops = 0.0
    for _ in range(4):   # the first probe can run while the clocks still ramp up: the best of four is the peak
        _lib.check(X.lib.skg_fp64_peak(4096, C.byref(pk_ms), C.byref(pk_fma)), "skg_fp64_peak")
        X.peak_tflops = max(X.peak_tflops, 2.0 * pk_fma.value / (pk_ms.value * 1e-3) / 1e12)

    sampler = ClockSampler(local_rank)
    sampler.start()
    import warnings
    warnings.simplefilter("ignore")

    # ---------------------------------------------------------------- headline: config 3, cressie
    c3, v3 = make_inputs(**C3)
    n3 = C3["n"]
    pairs3 = n3 * (n3 - 1) // 2
    V3 = skg.Variogram(c3, v3, estimator="cressie", n_lags=N_LAGS, maxlag="median", fit_method=None,
                       process_group=X.group)
    edges3 = V3.bins
    step_ms, kern_ms, info, launches, _ = fused_leg(X, c3, v3, "cressie", args.steps, args.warmup, edges3)
    step_ms, kern_ms = max_over_ranks(step_ms), max_over_ranks(kern_ms)
    e2e_steps = max(3, min(args.steps, 10))
    e2e_s, h2d, d2h, exp = e2e_leg(
        X, lambda: skg.Variogram(c3, v3, estimator="cressie", n_lags=N_LAGS, maxlag="median",
                                 fit_method=None, process_group=X.group).experimental, e2e_steps)
    assert np.allclose(exp, V3.experimental, rtol=1e-12)
    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": pairs3 / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": HEADLINE, "n_points": n3, "pairs": pairs3, "n_lags": N_LAGS,
                       "estimator": "cressie",
                       "parallelism": "tile triangle dealt round-robin over %d rank(s) + 1 all-reduce" % world,
                       "l2": "256 MiB buffer written between timed iterations (inputs are 4.8 MB and "
                             "cache-resident by design)"},
            "e2e": {"value": pairs3 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                    "call": "Variogram(coords, values, 'cressie', n_lags=25, maxlag='median').experimental"},
            "gpu_launches": int(launches),
            "roofline": roofline(X, FLOPS["cressie"], pairs3 / world, info, kern_ms,
                                 "pair_hist_win_kernel<2,1> (cressie: sum sqrt|dz|)", NCU["cressie"]),
        }
        out["roofline"]["hbm"] = {"algorithmic_bytes": 8.0 * 3 * n3 + 24.0 * N_LAGS,
                                  "achieved_gbs": (8.0 * 3 * n3 + 24.0 * N_LAGS) / (kern_ms * 1e-3) / 1e9}

    # ---------------------------------------------------------------- config 3, matheron and dowd
    sm, km, info_m, _, _ = fused_leg(X, c3, v3, "matheron", max(3, args.steps // 2), 3, edges3)
    sm, km = max_over_ranks(sm), max_over_ranks(km)
    em_s, _, _, _ = e2e_leg(
        X, lambda: skg.Variogram(c3, v3, estimator="matheron", n_lags=N_LAGS, maxlag="median",
                                 fit_method=None, process_group=X.group).experimental, 5, warm=3)
    ed_s, dh2d, dd2h, _ = e2e_leg(
        X, lambda: skg.Variogram(c3, v3, estimator="dowd", n_lags=N_LAGS, maxlag="median",
                                 fit_method=None, process_group=X.group).experimental, 5, warm=3)
    dk_ms, d_inside, d_pairs = dowd_sweep_leg(X, c3, v3, edges3, 3)
    dk_ms = max_over_ranks(dk_ms)
    if rank == 0:
        out["config3_matheron"] = {
            "value": pairs3 / (sm * 1e-3), "unit": UNIT, "ms_per_step": sm,
            "e2e": {"value": pairs3 / em_s, "unit": UNIT, "ms_per_step": em_s * 1e3},
            "roofline": roofline(X, FLOPS["matheron"], pairs3 / world, info_m, km,
                                 "pair_hist_win_kernel<1,1> (matheron: sum dz^2)", NCU["matheron"])}
        out["config3_dowd"] = {
            "e2e": {"value": pairs3 / ed_s, "unit": UNIT, "ms_per_step": ed_s * 1e3,
                    "h2d_bytes_per_step": int(dh2d), "d2h_bytes_per_step": int(dd2h),
                    "call": "Variogram(coords, values, 'dowd', n_lags=25, maxlag='median').experimental"},
            "roofline": {
                "bound": "fp64", "kernel": "pair_hist_win_kernel<4,1> (skg_pair_dz_windows: per-lag count, count "
                                           "below the median window, in-window |dz| histogram)",
                "kernel_ms": dk_ms, "flops_per_pair": FLOPS["dowd"],
                "achieved": FLOPS["dowd"] * d_pairs / (dk_ms * 1e-3) / 1e12, "peak": X.peak_tflops,
                "unit": "TFLOP/s", "frac": FLOPS["dowd"] * d_pairs / (dk_ms * 1e-3) / 1e12 / X.peak_tflops,
                "traffic": 6.7e6, "in_window_share": d_inside,
                "fp64_pipe_pct": 34.7, "issue_active_pct": 62.1,
                "from": "profiles/r2_pair_dz_windows_ncu_summary.txt (N=200k, 21.75 ms under ncu)",
                "note": "SURVEY 8(d): 8 ops per pair for the dowd sweep; compare-bound, not flop-bound: 6 FP64 arithmetic + 4 FP64 compares per pair, ~15 integer / "
                        "predicate instructions; achieved counts the in-range pairs (the per-lag counts) only; the second "
                        "kernel of the path is pair_dz_gather_kernel (8.8 ms, profiles/r2_pair_dz_gather_ncu_summary.txt)"},
            "note": "per-lag exact median of |dz|: sub-sample windows, one counting sweep, one gather"}

    # ---------------------------------------------------------------- config 2 (single-GPU quote)
    c2, v2 = make_inputs(**C2)
    pairs2 = C2["n"] * (C2["n"] - 1) // 2
    V2 = skg.Variogram(c2, v2, n_lags=N_LAGS, maxlag="median", fit_method=None, process_group=X.group)
    s2, k2, info2, l2, _ = fused_leg(X, c2, v2, "matheron", args.steps, args.warmup, V2.bins)
    s2, k2 = max_over_ranks(s2), max_over_ranks(k2)
    e2_s, h2, d2, exp2 = e2e_leg(
        X, lambda: skg.Variogram(c2, v2, n_lags=N_LAGS, maxlag="median", fit_method=None,
                                 process_group=X.group).experimental, max(5, args.steps), warm=5)
    assert np.allclose(exp2, V2.experimental, rtol=1e-12)
    if rank == 0:
        out["config2"] = {
            "workload": "2-D synthetic Gaussian field N=20000 (199990000 pairs), matheron, n_lags=25, "
                        "maxlag=median (BASELINE.json config 2)",
            "value": pairs2 / (s2 * 1e-3), "unit": UNIT, "ms_per_step": s2, "gpu_launches": int(l2),
            "e2e": {"value": pairs2 / e2_s, "unit": UNIT, "ms_per_step": e2_s * 1e3,
                    "h2d_bytes_per_step": int(h2), "d2h_bytes_per_step": int(d2)},
            "roofline": roofline(X, FLOPS["matheron"], pairs2 / world, info2, k2,
                                 "pair_hist_win_kernel<1,1> (per-pair path: windows off at this density)")}

    # ---------------------------------------------------------------- config 4 (directional, uniform)
    c4, v4 = make_inputs(**C4)
    pairs4 = C4["n"] * (C4["n"] - 1) // 2
    mk4 = lambda: skg.DirectionalVariogram(c4, v4, azimuth=45, tolerance=22.5, bin_func="uniform",
                                           n_lags=10, fit_method=None, process_group=X.group)
    D4 = mk4()
    _ = D4.experimental
    e4_s, h4, d4, _ = e2e_leg(X, lambda: mk4().experimental, 5, warm=3)
    s4, k4, info4, _, _ = fused_leg(X, c4, v4, "matheron", max(3, args.steps // 2), 3, D4.bins, D4._dirp())
    s4, k4 = max_over_ranks(s4), max_over_ranks(k4)
    if rank == 0:
        out["config4"] = {
            "workload": "DirectionalVariogram N=100000 (4999950000 pairs), azimuth=45, tolerance=22.5, "
                        "uniform bins, n_lags=10 (BASELINE.json config 4)",
            "value": pairs4 / (s4 * 1e-3), "unit": UNIT, "ms_per_step": s4,
            "e2e": {"value": pairs4 / e4_s, "unit": UNIT, "ms_per_step": e4_s * 1e3,
                    "h2d_bytes_per_step": int(h4), "d2h_bytes_per_step": int(d4)},
            "roofline": roofline(X, FLOPS["directional"], pairs4 / world, info4, k4,
                                 "pair_hist_kernel<2,1,1,...> (fused angular mask)")}

    krig = bench_kriging(skg, args, X, world, rank)
    batched = bench_batched(skg, V2, c2, v2) if world == 1 else None
    clocks = sampler.stop()
    if rank == 0:
        out["clocks"] = clocks
        out["kriging"] = krig
        if batched is not None:
            out["batched"] = batched
    if world == 1:
        # CPU baseline: the stock reference on a bounded sample (about 10-30 s of CPU work)
        rate, dt, kind, used = reference_variogram_rate(7000, "cressie")
        out["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": used, "kind": kind, "cpu_model": cpu_model(),
            "cpu_count": os.cpu_count(),
            "sample": "%s Variogram(coords, values, 'cressie', n_lags=25, maxlag='median') + .experimental "
                      "on N=7000 points (%d pairs), one process (the reference is single-threaded), %.1f s"
                      % ("stock reference" if kind == "reference" else "oracle port", 7000 * 6999 // 2, dt)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_batched(skg, V, c, v, k=32):
    """k value vectors over the same coordinates and bins (Monte-Carlo
    propagation) through the batched pair sweep, next to one fused pass per vector.  Warmed with
    the same shape; median of 5 timed calls."""
    import torch
    rng = np.random.default_rng(1)
    draws = np.vstack([v + rng.normal(0.0, 0.05 * np.std(v), v.size) for _ in range(k)])
    pairs = v.size * (v.size - 1) // 2
    for _ in range(2):
        V.experimental_batch(draws)
    torch.cuda.synchronize()
    tbs = []
    for _ in range(5):
        t0 = time.perf_counter()
        got = V.experimental_batch(draws)
        torch.cuda.synchronize()
        tbs.append(time.perf_counter() - t0)
    tb = float(np.median(tbs))
    edges = V.bins
    t0 = time.perf_counter()
    for q in range(k):
        one = skg.Variogram(V._X, draws[q], n_lags=len(edges), maxlag=V.maxlag, fit_method=None).experimental
    torch.cuda.synchronize()
    ts = time.perf_counter() - t0
    assert np.allclose(got[-1], one, rtol=1e-12)
    V.experimental  # leaves the engine's values as they were
    return {"metric": "vector-pairs/s", "vectors": k,
            "e2e": {"value": k * pairs / tb, "unit": "vector-pairs/s", "ms": tb * 1e3,
                    "call": "Variogram.experimental_batch(values[%d, N]), median of 5 warmed calls" % k},
            "per_vector_e2e": {"value": k * pairs / ts, "unit": "vector-pairs/s", "ms": ts * 1e3,
                               "call": "Variogram(MetricSpace, values_k).experimental x %d" % k}}


def bench_kriging(skg, args, X, world, rank):
    """BASELINE.json config 5: 2000x2000 grid from N=10k samples, min 5 / max 64, exponential.  With
    N ranks the targets are split into contiguous slices, no collective (strong scaling)."""
    import torch
    from skgstat_b200.engine import Traffic, to_device
    c, v = make_inputs(10000, 1000.0, 100.0, 5)
    V = skg.Variogram(c, v, model="exponential", n_lags=25, maxlag="median")
    ok = skg.OrdinaryKriging(V, min_points=5, max_points=64, process_group=X.group)
    side = args.krige_side
    gx, gy = np.mgrid[0:1000:complex(0, side), 0:1000:complex(0, side)].reshape(2, -1)
    m = gx.size
    eng = ok._engine()
    tg = to_device(np.vstack((gx, gy)), eng.device)
    mp = ok._model_params()
    eng.transform_device(tg, m, "exponential", mp, 5, 64)
    X.barrier()
    evs = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        z, s, st = eng.transform_device(tg, m, "exponential", mp, 5, 64)
        e1.record()
        evs.append((e0, e1))
    X.barrier()
    ms = X.max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in evs])))
    # Cholesky path actually executed per target with n neighbours: n^3/3 (factor)
    # + 2 n^2 (two right-hand sides) + n^2 (back substitution) + 6 n(n-1)/2 (distances) + ~25 per
    # model evaluation; the reference's LU formulation of the same system would be 2.6e5
    n_nb = 64
    flops_exec = n_nb ** 3 / 3 + 3 * n_nb ** 2 + 3 * n_nb * (n_nb - 1) + 25 * (n_nb * (n_nb + 1) // 2)
    out = {"metric": "kriged points/s", "value": m / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms,
           "targets": int(m), "samples": 10000, "min_points": 5, "max_points": 64, "model": "exponential",
           "scaling": "strong", "parallelism": "target slices x%d, no collective" % world,
           "roofline": {"bound": "fp64", "unit": "TFLOP/s", "peak": X.peak_tflops,
                        "flops_per_point_executed": flops_exec,
                        "achieved": flops_exec * m / (ms * 1e-3) / 1e12,
                        "frac": flops_exec * m / (ms * 1e-3) / 1e12 / X.peak_tflops,
                        "flops_per_point_reference_lu": 2.6e5,
                        "achieved_counting_reference_lu": 2.6e5 * m / (ms * 1e-3) / 1e12,
                        "note": "the kernel solves the covariance form by Cholesky (about 1/2 of the "
                                "reference's LU flops); `achieved`/`frac` count the executed formulation",
                        "kernel": "krige_chol_kernel<2,2>", "kernel_ms": ms,
                        "traffic": NCU["kriging"]["traffic_bytes"] * m / 160000.0,
                        "fp64_pipe_pct": NCU["kriging"]["fp64_pipe_pct"],
                        "issue_active_pct": NCU["kriging"]["issue_active_pct"],
                        "smem_wavefronts_pct": NCU["kriging"]["smem_wavefronts_pct"],
                        "ncu_from": NCU["kriging"]["from"] + "; traffic scaled by the target count"}}
    if world == 1:
        ok.transform(gx, gy)   # warm-up of the public call (pinned staging buffers are allocated once)
        Traffic.reset()
        t0 = time.perf_counter()
        zz = ok.transform(gx, gy)
        e2e_s = time.perf_counter() - t0
        out["nan_fraction"] = float(np.mean(np.isnan(zz)))
        out["e2e"] = {"value": m / e2e_s, "unit": "points/s", "h2d_bytes_per_step": Traffic.h2d,
                      "d2h_bytes_per_step": Traffic.d2h, "call": "OrdinaryKriging.transform(gx, gy)"}
        ref = reference_kriging_rate(1500) if side == 2000 else None
        if ref is not None:
            base, sub, zr = ref
            out["cpu_baseline"] = base
            assert np.allclose(zz[sub], zr, rtol=1e-9, equal_nan=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--krige-side", type=int, default=2000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
