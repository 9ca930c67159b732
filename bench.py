#!/usr/bin/env python
"""bench.py - variogram pairs/s (fused distance + lag bin + estimator) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm (default).  One "step" = one fused pass (skg_pair_hist: distances, lag
classes, per-lag count and sum dz^2) over every pair of the point set, inputs
resident in HBM.  N=1 runs BASELINE.json config 2: 2-D synthetic Gaussian field,
20 000 points (199 990 000 pairs), matheron, n_lags=25, maxlag='median'.  For
N>1 (torchrun, one rank per GPU, NCCL) the point count grows as
20000*sqrt(N) so that every GPU keeps ~2.0e8 pairs (weak scaling); each rank
sweeps its tile range and the per-lag partials meet in ONE all-reduce per step.
`e2e` is the same metric through the public class API with HOST buffers:
Variogram(coords, values, 'matheron', n_lags=25, maxlag='median') - upload,
exact median select, edges, fused pass, download - per step.
Extra keys: `roofline` (FP64-ALU bound; denominator measured here with a DFMA
chain), `cpu_baseline` (oracle port on the host, bounded sample), `kriging`
(BASELINE.json config 5: kriged points/s, same line for convenience).

Reference arm (--impl reference): the reference algorithm's CPU path - the
numpy/scipy oracle port (oracle/variogram.py: pdist + digitize + per-lag sweeps,
single-threaded like the reference) - on a bounded sample of the same workload.
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
N_POINTS = 20000
N_LAGS = 25
FLOPS_PER_PAIR = 9.0  # 3 sub + 3 mul + 2 add + 1 sqrt-equivalent (2-D, matheron)


def make_inputs(n, seed=0):
    from skgstat_b200.data import gaussian_field, random_points
    c = random_points(n, 1000.0, 2, seed=seed)
    v = gaussian_field(c, 100.0, seed=seed)
    return c, v


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

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


def cpu_oracle_pairs_per_s(n_sample, repeats=1):
    """The reference algorithm on the host (oracle port; 1 thread like the reference)."""
    from oracle import variogram as ov
    c, v = make_inputs(n_sample, seed=1)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        ov.dense_variogram(c, v, "matheron", "even", N_LAGS, "median")
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    pairs = n_sample * (n_sample - 1) // 2
    return pairs / best, best, pairs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = 8000
    times = []
    for i in range(args.warmup + args.steps):
        rate, dt, pairs = cpu_oracle_pairs_per_s(n_sample)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = pairs / (ms * 1e-3)
    sample = "N=%d points (%d pairs) of the N=%d workload per step" % (n_sample, pairs, N_POINTS)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "2-D synthetic Gaussian field N=20000, matheron, n_lags=25, "
                               "maxlag=median (BASELINE.json config 2)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_own(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import skgstat_b200 as skg
    from skgstat_b200 import _lib
    from skgstat_b200.engine import PairEngine, Traffic

    lib = _lib.load()
    n = int(round(N_POINTS * np.sqrt(world)))
    c, v = make_inputs(n)
    pairs = n * (n - 1) // 2

    # edges once, through the public path (exact median select)
    V = skg.Variogram(c, v, n_lags=N_LAGS, maxlag="median", fit_method=None)
    edges = V.bins
    eng = PairEngine(c, v, group=group)
    from skgstat_b200._exact import sq_threshold_bits
    thr = sq_threshold_bits(edges)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        eng.hist(edges, _lib.STAT_SUM_SQ, None, thr_bits=thr)
    barrier()
    l0 = eng.launches
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cnt, ssq, _ = eng.hist(edges, _lib.STAT_SUM_SQ, None, thr_bits=thr)
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = eng.launches - l0
    step_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    assert int(cnt.sum()) > 0 and np.all(np.isfinite(ssq))

    # kernel-only duration of the dominant kernel (skg_pair_hist incl. its tiny reduce),
    # CUDA events directly around the C-ABI call on the launching stream
    import ctypes as C
    from skgstat_b200.engine import _hptr, _ptr
    nl = N_LAGS
    cnt_d = torch.zeros(nl, dtype=torch.int64, device="cuda")
    sq_d = torch.zeros(nl, dtype=torch.float64, device="cuda")
    scratch = eng._hist_scratch(nl)
    kev = []
    for i in range(3 + args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.skg_pair_hist(_ptr(eng.coords), _ptr(eng.values), 1, 0, eng.n, eng.dim, None, _hptr(thr),
                               nl, _lib.STAT_SUM_SQ, None, _ptr(cnt_d), _ptr(sq_d), None, _ptr(scratch),
                               scratch.numel(), eng.t0, eng.t1, eng.tstride, eng._stream())
        e1.record()
        _lib.check(rc, "skg_pair_hist")
        if i >= 3:
            kev.append((e0, e1))
    barrier()
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    kept, cand, slots = eng.last_sweep_info()  # tile culling: work items swept / before culling

    # end to end through the class API with host buffers
    e2e_steps = max(5, args.steps)
    for _ in range(5):
        skg.Variogram(c, v, n_lags=N_LAGS, maxlag="median", fit_method=None,
                      process_group=group).experimental
    barrier()
    Traffic.reset()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        Ve = skg.Variogram(c.copy(), v.copy(), n_lags=N_LAGS, maxlag="median", fit_method=None,
                           process_group=group)
        exp = Ve.experimental
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    h2d, d2h = Traffic.h2d // e2e_steps, Traffic.d2h // e2e_steps
    assert np.allclose(exp, V.experimental, rtol=1e-12)
    krig = None
    batched = None
    if world == 1 and not args.no_kriging:
        batched = bench_batched(skg, V, c, v)
        krig = bench_kriging(skg, args)  # sampled too: the longest loaded stretch of the run
    elif not args.no_kriging:
        krig = bench_kriging_sharded(skg, args, group, world)
    clocks = sampler.stop()

    t = torch.tensor([step_ms, kern_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms, e2e_s = [float(x) for x in t.cpu()]

    out = None
    if rank == 0:
        pk_ms, pk_fma = C.c_double(), C.c_double()
        _lib.check(lib.skg_fp64_peak(4096, C.byref(pk_ms), C.byref(pk_fma)), "skg_fp64_peak")
        peak_tflops = 2.0 * pk_fma.value / (pk_ms.value * 1e-3) / 1e12
        my_pairs = pairs / world  # the dominant kernel of one rank covers its share of the tiles
        # hardware efficiency is judged on the pairs the kernel actually evaluates: tiles
        # beyond maxlag are culled from bounding boxes before the sweep (exact)
        evaluated = min(float(kept) * slots, my_pairs) if cand else my_pairs
        achieved = FLOPS_PER_PAIR * evaluated / (kern_ms * 1e-3) / 1e12
        alg_bytes = 8.0 * 3 * n + 24.0 * N_LAGS
        out = {
            "metric": METRIC, "value": pairs / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "2-D synthetic Gaussian field N=%d (%d pairs), matheron, "
                                   "n_lags=25, maxlag=median (BASELINE.json config 2%s)"
                                   % (n, pairs, "" if world == 1 else
                                      "; points scaled by sqrt(gpus) to keep pairs/GPU fixed"),
                       "n_points": n, "pairs": pairs, "n_lags": N_LAGS,
                       "parallelism": "tile-range x%d + 1 all-reduce" % world,
                       "l2": "256 MiB buffer written between timed iterations (inputs are "
                             "0.5 MB and cache-resident by design)"},
            "clocks": clocks,
            "e2e": {"value": pairs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                    "call": "Variogram(coords, values, 'matheron', n_lags=25, maxlag='median')"
                            ".experimental"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tflops,
                         "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this
                         # config, from the committed ncu --set full capture
                         # (profiles/r1_pair_hist_ncu_summary.txt: 590 592 B read, 0 written); algorithmic bytes: 480 600
                         "traffic": 590592 if world == 1 else None,
                         "kernel": "pair_hist_kernel<2,0,1,1,true>", "kernel_ms": kern_ms,
                         "pairs_evaluated": evaluated, "pairs_total": my_pairs,
                         "culled_fraction": 1.0 - evaluated / my_pairs,
                         "achieved_counting_culled_pairs": FLOPS_PER_PAIR * my_pairs
                         / (kern_ms * 1e-3) / 1e12,
                         "flops_per_pair": FLOPS_PER_PAIR,
                         "peak_source": "DFMA chain measured in this run (skg_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 figure",
                         "hbm": {"achieved_gbs": alg_bytes / (kern_ms * 1e-3) / 1e9,
                                 "algorithmic_bytes": alg_bytes}},
        }
    if world == 1:
        # CPU baseline (oracle port, bounded sample) and the kriging metric, rank 0 / N=1 only
        rate, dt, sp = cpu_oracle_pairs_per_s(6000)
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": "oracle.dense_variogram on N=6000 points (%d pairs), "
                                         "%.1f s" % (sp, dt)}
        if krig is not None:
            out["kriging"] = krig
        if batched is not None:
            out["batched"] = batched
    if world > 1 and rank == 0 and krig is not None:
        out["kriging"] = krig
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_batched(skg, V, c, v, k=32):
    """k value vectors over the same coordinates and bins (Monte-Carlo
    propagation) through the batched pair sweep, next to one fused pass per vector."""
    import torch
    rng = np.random.default_rng(1)
    draws = np.vstack([v + rng.normal(0.0, 0.05 * np.std(v), v.size) for _ in range(k)])
    pairs = v.size * (v.size - 1) // 2
    V.experimental_batch(draws[:4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = V.experimental_batch(draws)
    torch.cuda.synchronize()
    tb = time.perf_counter() - t0
    edges = V.bins
    t0 = time.perf_counter()
    for q in range(k):
        one = skg.Variogram(V._X, draws[q], n_lags=len(edges), maxlag=V.maxlag, fit_method=None).experimental
    torch.cuda.synchronize()
    ts = time.perf_counter() - t0
    assert np.allclose(got[-1], one, rtol=1e-12)
    V.experimental  # leaves the engine's values as they were
    return {"metric": "vector-pairs/s", "vectors": k, "e2e": {"value": k * pairs / tb, "unit": "vector-pairs/s",
            "ms": tb * 1e3, "call": "Variogram.experimental_batch(values[%d, N])" % k},
            "per_vector_e2e": {"value": k * pairs / ts, "unit": "vector-pairs/s", "ms": ts * 1e3,
                               "call": "Variogram(MetricSpace, values_k).experimental x %d" % k}}


def bench_kriging_sharded(skg, args, group, world):
    """Config 5 on N GPUs: the 4e6 targets are split into contiguous slices, one per rank, no
    collective on the data path (strong scaling of the kriging metric); device-timed, max over
    ranks."""
    import torch
    import torch.distributed as dist
    from skgstat_b200.engine import to_device
    c, v = make_inputs(10000, seed=5)
    V = skg.Variogram(c, v, model="exponential", n_lags=25, maxlag="median")
    ok = skg.OrdinaryKriging(V, min_points=5, max_points=64, process_group=group)
    side = args.krige_side
    gx, gy = np.mgrid[0:1000:complex(0, side), 0:1000:complex(0, side)].reshape(2, -1)
    m = gx.size
    eng = ok._engine()
    tg = to_device(np.vstack((gx, gy)), eng.device)
    mp = ok._model_params()
    eng.transform_device(tg, m, "exponential", mp, 5, 64)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        eng.transform_device(tg, m, "exponential", mp, 5, 64)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 2], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"metric": "kriged points/s", "value": m / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms,
            "targets": int(m), "samples": 10000, "min_points": 5, "max_points": 64, "model": "exponential",
            "scaling": "strong", "parallelism": "target slices x%d, no collective" % world}


def bench_kriging(skg, args):
    """BASELINE.json config 5: 2000x2000 grid from N=10k samples, min 5 / max 64, exponential."""
    import torch
    from skgstat_b200.engine import Traffic
    c, v = make_inputs(10000, seed=5)
    V = skg.Variogram(c, v, model="exponential", n_lags=25, maxlag="median")
    ok = skg.OrdinaryKriging(V, min_points=5, max_points=64)
    side = args.krige_side
    gx, gy = np.mgrid[0:1000:complex(0, side), 0:1000:complex(0, side)].reshape(2, -1)
    m = gx.size
    eng = ok._engine()
    from skgstat_b200.engine import to_device
    tg = to_device(np.vstack((gx, gy)), eng.device)
    mp = ok._model_params()
    for _ in range(1):
        eng.transform_device(tg, m, "exponential", mp, 5, 64)
    torch.cuda.synchronize()
    evs = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        z, s, st = eng.transform_device(tg, m, "exponential", mp, 5, 64)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    ok.transform(gx, gy)   # warm-up of the public call (pinned staging buffers are allocated once)
    Traffic.reset()
    t0 = time.perf_counter()
    zz = ok.transform(gx, gy)
    e2e_s = time.perf_counter() - t0
    d = V.describe()
    # CPU baseline: the reference algorithm (oracle port, 1 core) on a bounded sample
    from oracle import kriging as okr
    sub = np.random.default_rng(0).choice(m, 400, replace=False)
    t0 = time.perf_counter()
    zr, sr, _ = okr.krige(c, v, np.column_stack((gx[sub], gy[sub])), "exponential",
                          d["effective_range"], d["sill"], d["nugget"], None, 5, 64)
    cpu_s = time.perf_counter() - t0
    assert np.allclose(zz[sub], zr, rtol=1e-9, equal_nan=True)
    return {"metric": "kriged points/s", "value": m / (ms * 1e-3), "unit": "points/s",
            "cpu_baseline": {"value": 400 / cpu_s, "unit": "points/s", "cores": 1, "kind": "port",
                             "sample": "oracle.krige on 400 random grid nodes, %.1f s" % cpu_s},
            "ms_per_step": ms, "targets": int(m), "samples": 10000, "min_points": 5,
            "max_points": 64, "model": "exponential", "effective_range": d["effective_range"],
            "nan_fraction": float(np.mean(np.isnan(zz))),
            "e2e": {"value": m / e2e_s, "unit": "points/s", "h2d_bytes_per_step": Traffic.h2d,
                    "d2h_bytes_per_step": Traffic.d2h},
            "flops_per_point": 2.6e5,
            "achieved_tflops": 2.6e5 * m / (ms * 1e-3) / 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-kriging", action="store_true")
    ap.add_argument("--krige-side", type=int, default=2000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
