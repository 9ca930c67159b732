"""Test infrastructure (not product): generate tests/golden/*.npz by running
the UNMODIFIED reference (scikit-gstat 1.0.23 imported from /root/reference,
see oracle/ref_harness.py) in the build container.

    python oracle/make_golden.py

The GPU box has no /root/reference, so the vectors are committed.  Every
array is the reference's own output; inputs are stored next to the outputs so
the parity tests need nothing else.  Case parameters live in a JSON manifest
inside each npz (key ``manifest``).
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.fields import gaussian_field, random_points  # noqa: E402
from oracle.ref_harness import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def datasets(skg):
    ds = {}
    c, v = skg.data.pancake(N=300, seed=42).get("sample")
    ds["pancake300"] = (np.asarray(c), np.asarray(v))
    c = random_points(400, 100.0, 2, seed=1)
    ds["rnd400"] = (c, gaussian_field(c, 12.0, seed=1))
    gx, gy = np.meshgrid(np.arange(15), np.arange(15))
    c = np.column_stack((gx.ravel(), gy.ravel())).astype(float)
    ds["lattice225"] = (c, gaussian_field(c, 3.0, seed=2))
    c = random_points(150, 50.0, 3, seed=3)
    ds["rnd3d150"] = (c, gaussian_field(c, 8.0, seed=3))
    c = random_points(120, 80.0, 1, seed=4)[:, 0]
    ds["rnd1d120"] = (c, gaussian_field(c[:, None], 6.0, seed=4))
    # the reference's own unit-test inputs (tests/test_variogram.py:72-79)
    np.random.seed(42)
    c = np.random.gamma(10, 4, (30, 2))
    np.random.seed(42)
    v = np.random.normal(10, 4, 30)
    ds["gamma30"] = (c, v)
    return ds


def make_variogram(skg, ds):
    arrays, manifest = {}, []
    lag_cfgs = [(10, None), (25, "median"), (8, "mean"), (6, 0.6), (12, "abs")]
    for dname, (c, v) in ds.items():
        arrays["in/%s/coords" % dname] = c
        arrays["in/%s/values" % dname] = v
        for est in ("matheron", "cressie", "dowd"):
            for bf in ("even", "uniform"):
                for n_lags, ml in lag_cfgs:
                    if ml == "abs":
                        from scipy.spatial.distance import pdist
                        cc = c if c.ndim == 2 else np.column_stack((c, np.zeros(len(c))))
                        ml_val = float(np.round(0.45 * pdist(cc).max(), 3)) + 1.0
                    else:
                        ml_val = ml
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        V = skg.Variogram(c, v, estimator=est, bin_func=bf,
                                          n_lags=n_lags, maxlag=ml_val)
                    key = "v/%d" % len(manifest)
                    arrays[key + "/bins"] = V.bins
                    arrays[key + "/bin_count"] = np.asarray(V.bin_count, dtype=np.int64)
                    arrays[key + "/experimental"] = V.experimental
                    arrays[key + "/maxlag"] = np.array(
                        np.nan if V.maxlag is None else V.maxlag, dtype=float)
                    arrays[key + "/cof"] = np.asarray(V.cof, dtype=float)
                    manifest.append(dict(key=key, dataset=dname, estimator=est,
                                         bin_func=bf, n_lags=n_lags, maxlag=ml_val))
    arrays["manifest"] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(OUT, "variogram.npz"), **arrays)
    print("variogram cases:", len(manifest))


def make_directional(skg, ds):
    arrays, manifest = {}, []
    dir_cfgs = [
        dict(azimuth=0, tolerance=45.0, bandwidth="q33", directional_model="triangle"),
        dict(azimuth=45, tolerance=22.5, bandwidth="q33", directional_model="triangle"),
        dict(azimuth=-120, tolerance=10.0, bandwidth=30.0, directional_model="triangle"),
        dict(azimuth=90, tolerance=90.0, bandwidth="q33", directional_model="compass"),
        dict(azimuth=45, tolerance=22.5, bandwidth="q33", directional_model="compass"),
        dict(azimuth=170, tolerance=60.0, bandwidth="q50", directional_model="triangle"),
    ]
    for dname in ("pancake300", "rnd400", "gamma30"):
        c, v = ds[dname]
        arrays["in/%s/coords" % dname] = c
        arrays["in/%s/values" % dname] = v
        for cfg in dir_cfgs:
            for bf, est, n_lags, ml in (("even", "matheron", 10, None),
                                        ("uniform", "matheron", 10, None),
                                        ("even", "cressie", 8, "median"),
                                        ("uniform", "dowd", 6, "median")):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    V = skg.DirectionalVariogram(c, v, estimator=est, bin_func=bf,
                                                 n_lags=n_lags, maxlag=ml, **cfg)
                key = "d/%d" % len(manifest)
                arrays[key + "/bins"] = V.bins
                arrays[key + "/bin_count"] = np.asarray(V.bin_count, dtype=np.int64)
                arrays[key + "/experimental"] = V.experimental
                arrays[key + "/maxlag"] = np.array(
                    np.nan if V.maxlag is None else V.maxlag, dtype=float)
                arrays[key + "/bandwidth"] = np.array(V.bandwidth, dtype=float)
                arrays[key + "/mask_count"] = np.array(int(V._direction_mask().sum()))
                manifest.append(dict(key=key, dataset=dname, estimator=est, bin_func=bf,
                                     n_lags=n_lags, maxlag=ml, **cfg))
    arrays["manifest"] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(OUT, "directional.npz"), **arrays)
    print("directional cases:", len(manifest))


def make_kriging(skg, ds):
    arrays, manifest = {}, []
    rng = np.random.default_rng(11)
    c1000 = random_points(1000, 500.0, 2, seed=5)
    local = dict(ds)
    local["rnd1000"] = (c1000, gaussian_field(c1000, 60.0, seed=5))
    c3 = random_points(400, 60.0, 3, seed=6)
    local["rnd3d400"] = (c3, gaussian_field(c3, 12.0, seed=6))
    cases = [
        dict(dataset="pancake300", model="exponential", n_lags=25, maxlag="median",
             min_points=5, max_points=64, targets="grid6"),
        dict(dataset="pancake300", model="spherical", n_lags=10, maxlag=None,
             min_points=5, max_points=15, targets="rand120+samples"),
        dict(dataset="rnd1000", model="exponential", n_lags=25, maxlag="median",
             min_points=5, max_points=64, targets="rand300"),
        dict(dataset="rnd1000", model="spherical", n_lags=20, maxlag=0.3,
             min_points=8, max_points=20, targets="rand300far"),
        dict(dataset="rnd1000", model="cubic", n_lags=20, maxlag="median",
             min_points=3, max_points=12, targets="rand200"),
        # stable with a manual shape: the free fit runs into shape=2 (Gaussian-like,
        # cond(A) ~ 1e18, no meaningful parity - SURVEY.md section 7.3 item 4)
        dict(dataset="rnd1000", model="stable", n_lags=20, maxlag="median",
             min_points=3, max_points=24, targets="rand200",
             manual=dict(fit_range=150.0, fit_sill=3.5, fit_shape=1.3)),
        dict(dataset="rnd1000", model="exponential", n_lags=25, maxlag="median",
             min_points=5, max_points=32, targets="rand200", use_nugget=True),
        dict(dataset="rnd3d400", model="exponential", n_lags=15, maxlag="median",
             min_points=4, max_points=16, targets="rand150"),
        dict(dataset="lattice225", model="spherical", n_lags=8, maxlag=None,
             min_points=4, max_points=10, targets="latticehalf"),
    ]
    for case in cases:
        c, v = local[case["dataset"]]
        arrays["in/%s/coords" % case["dataset"]] = c
        arrays["in/%s/values" % case["dataset"]] = v
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            extra = {}
            if "manual" in case:
                extra = dict(fit_method="manual", **case["manual"])
            V = skg.Variogram(c, v, model=case["model"], n_lags=case["n_lags"],
                              maxlag=case["maxlag"],
                              use_nugget=case.get("use_nugget", False), **extra)
        lo, hi = c.min(axis=0), c.max(axis=0)
        dim = c.shape[1]
        kind = case["targets"]
        if kind == "grid6":
            t = np.mgrid[0:499:6j, 0:499:6j].reshape(2, -1).T
        elif kind == "rand120+samples":
            t = np.vstack((lo + rng.random((120, dim)) * (hi - lo), c[:15].astype(float)))
        elif kind == "rand300far":
            t = lo - 0.4 * (hi - lo) + rng.random((300, dim)) * 1.8 * (hi - lo)
        elif kind == "latticehalf":
            gx, gy = np.meshgrid(np.arange(-1, 16, 0.5), np.arange(-1, 16, 0.5))
            t = np.column_stack((gx.ravel(), gy.ravel()))
        else:
            t = lo + rng.random((int(kind[4:]), dim)) * (hi - lo)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ok = skg.OrdinaryKriging(V, min_points=case["min_points"],
                                     max_points=case["max_points"])
            z = ok.transform(*[t[:, k] for k in range(dim)])
        d = V.describe()
        key = "k/%d" % len(manifest)
        arrays[key + "/targets"] = t
        arrays[key + "/z"] = z
        arrays[key + "/sigma"] = ok.sigma
        entry = dict(case)
        entry.update(key=key, effective_range=float(d["effective_range"]),
                     sill=float(d["sill"]), nugget=float(d["nugget"]),
                     shape=(float(d["shape"]) if "shape" in d else None),
                     no_points_error=int(ok.no_points_error),
                     singular_error=int(ok.singular_error))
        manifest.append(entry)
        print(key, case["dataset"], case["model"], "nan:", int(np.isnan(z).sum()), "of", len(z))
    arrays["manifest"] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(OUT, "kriging.npz"), **arrays)


def make_cross_validation(skg, ds):
    """Reference jacknife (util/cross_validation.py) deviations and metrics."""
    from skgstat.util import cross_validation as rcv
    from itertools import cycle
    arrays, manifest = {}, []
    c1000 = random_points(600, 500.0, 2, seed=5)
    local = dict(ds)
    local["rnd600"] = (c1000, gaussian_field(c1000, 60.0, seed=5))
    for dname, model, n, seed in (("pancake300", "spherical", None, 42),
                                  ("rnd600", "exponential", 250, 7),
                                  ("rnd400", "spherical", None, 1)):
        c, v = local[dname]
        arrays["in/%s/coords" % dname] = c
        arrays["in/%s/values" % dname] = v
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            V = skg.Variogram(c, v, model=model, n_lags=15, maxlag="median")
            rng = np.random.default_rng(seed=seed)
            size = n if n is not None else len(c)
            idx = rng.choice(len(c), replace=False, size=size)
            dev = np.fromiter(map(rcv._interpolate, idx, cycle([V])), dtype=float)
            metrics = [V.cross_validate(n=n, metric=m, seed=seed) for m in ("rmse", "mse", "mae")]
        d = V.describe()
        key = "cv/%d" % len(manifest)
        arrays[key + "/indices"] = idx
        arrays[key + "/deviations"] = dev
        arrays[key + "/metrics"] = np.asarray(metrics, dtype=float)
        manifest.append(dict(key=key, dataset=dname, model=model, n=n, seed=seed,
                             effective_range=float(d["effective_range"]), sill=float(d["sill"]),
                             nugget=float(d["nugget"])))
        print(key, dname, "rmse", metrics[0], "nan", int(np.isnan(dev).sum()))
    arrays["manifest"] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(OUT, "cross_validation.npz"), **arrays)


def make_multivariate(skg, ds):
    """Cross-variograms (|dz_0|*|dz_1|), multivariate='aggregate', the cross_variograms grid and
    the Monte-Carlo propagate utility - Variogram.py:1965-1984, util/cross_variogram.py:61-87,
    util/uncertainty.py:119-195."""
    from skgstat.util.cross_variogram import cross_variograms
    from skgstat.util.uncertainty import propagate
    arrays, manifest = {}, []
    mv_sets = {}
    c, v = ds["rnd400"]
    rng = np.random.default_rng(11)
    v2 = 0.6 * v + 0.8 * gaussian_field(c, 20.0, seed=21) + 0.05 * rng.normal(size=len(v))
    v3 = gaussian_field(c, 6.0, seed=22) + 3.0
    mv_sets["rnd400"] = (c, np.column_stack((v, v2, v3)))
    # the reference's own cross-variogram test input (tests/test_variogram.py:1599-1602)
    np.random.seed(42)
    cg = np.random.gamma(10, 4, (100, 2))
    np.random.seed(42)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vg = np.random.multivariate_normal([10.5, 5.6], [[1.5, 3.2], [3.2, 1.5]], size=100)
    mv_sets["gamma100"] = (cg, vg)
    c3, v3d = ds["rnd3d150"]
    mv_sets["rnd3d150"] = (c3, np.column_stack((v3d, gaussian_field(c3, 15.0, seed=23))))
    for dname, (c, vals) in mv_sets.items():
        arrays["in/%s/coords" % dname] = c
        arrays["in/%s/values" % dname] = vals
        for mode in ("cross", "aggregate"):
            cols = vals[:, :2] if mode == "cross" else vals
            for est in ("matheron", "cressie", "dowd"):
                for bf, n_lags, ml in (("even", 10, None), ("uniform", 8, "median"), ("even", 25, "median")):
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        V = skg.Variogram(c, cols, estimator=est, bin_func=bf, n_lags=n_lags,
                                          maxlag=ml, multivariate=mode)
                    key = "m/%d" % len(manifest)
                    arrays[key + "/bins"] = V.bins
                    arrays[key + "/bin_count"] = np.asarray(V.bin_count, dtype=np.int64)
                    arrays[key + "/experimental"] = V.experimental
                    arrays[key + "/diff_checksum"] = np.array([V.pairwise_diffs.sum(), V.pairwise_diffs.max()])
                    manifest.append(dict(key=key, kind="variogram", dataset=dname, mode=mode,
                                         estimator=est, bin_func=bf, n_lags=n_lags, maxlag=ml))
    # directional cross-variograms
    c, vals = mv_sets["rnd400"]
    for cfg in (dict(azimuth=45, tolerance=22.5, directional_model="triangle"),
                dict(azimuth=90, tolerance=60.0, directional_model="compass")):
        for est in ("matheron", "dowd"):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                V = skg.DirectionalVariogram(c, vals[:, :2], estimator=est, n_lags=8, maxlag="median", **cfg)
            key = "m/%d" % len(manifest)
            arrays[key + "/bins"] = V.bins
            arrays[key + "/bin_count"] = np.asarray(V.bin_count, dtype=np.int64)
            arrays[key + "/experimental"] = V.experimental
            manifest.append(dict(key=key, kind="directional", dataset="rnd400", mode="cross",
                                 estimator=est, n_lags=8, maxlag="median", **cfg))
    # cross_variograms grid: 3 variables -> 3x3 matrix
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mat = cross_variograms(c, vals, n_lags=12, maxlag="median")
    key = "m/%d" % len(manifest)
    arrays[key + "/experimental"] = np.array([[mat[i][j].experimental for j in range(3)] for i in range(3)])
    arrays[key + "/bin_count"] = np.array([[mat[i][j].bin_count for j in range(3)] for i in range(3)], dtype=np.int64)
    arrays[key + "/is_cross"] = np.array([[bool(mat[i][j].is_cross_variogram) for j in range(3)] for i in range(3)])
    manifest.append(dict(key=key, kind="grid", dataset="rnd400", n_lags=12, maxlag="median"))
    # Monte-Carlo uncertainty propagation (seeded)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        V = skg.Variogram(c, vals[:, 0], n_lags=10, maxlag="median", model="exponential")
        for est in ("matheron", "cressie"):
            V.estimator = est
            ci = propagate(V, "values", sigma=0.3, evalf="experimental", num_iter=40, seed=1234)
            key = "m/%d" % len(manifest)
            arrays[key + "/conf"] = np.asarray(ci)
            manifest.append(dict(key=key, kind="propagate", dataset="rnd400", estimator=est,
                                 sigma=0.3, num_iter=40, seed=1234, n_lags=10, maxlag="median"))
        V.estimator = "matheron"
        res = propagate(V, "values", sigma=0.2, evalf=["experimental", "parameter", "model"],
                        num_iter=12, seed=7, q=20, eval_at=30)
        key = "m/%d" % len(manifest)
        for nm, r in zip(("conf_exp", "conf_par", "conf_model"), res):
            arrays[key + "/" + nm] = np.asarray(r)
        manifest.append(dict(key=key, kind="propagate3", dataset="rnd400", sigma=0.2, num_iter=12,
                             seed=7, q=20, eval_at=30, n_lags=10, maxlag="median"))
    arrays["manifest"] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(OUT, "multivariate.npz"), **arrays)
    print("multivariate cases:", len(manifest))


def make_kriging_ext(skg, ds):
    """SURVEY.md 8(f) N2 leftovers: Matern model (models.py:347-422) and
    OrdinaryKriging(mode='estimate') (Kriging.py:343-352, 656-679)."""
    arrays, manifest = {}, []
    rng = np.random.default_rng(17)
    c = random_points(1000, 500.0, 2, seed=5)
    v = gaussian_field(c, 60.0, seed=5)
    arrays["in/rnd1000/coords"] = c
    arrays["in/rnd1000/values"] = v
    lo, hi = c.min(axis=0), c.max(axis=0)
    t = lo + rng.random((250, 2)) * (hi - lo)
    t = np.vstack((t, c[:10]))
    arrays["targets"] = t
    cases = [
        dict(model="matern", mode="exact", max_points=16, manual=dict(fit_range=180.0, fit_sill=3.4, fit_shape=0.7)),
        dict(model="matern", mode="exact", max_points=24, manual=dict(fit_range=200.0, fit_sill=3.4, fit_shape=1.5)),
        dict(model="matern", mode="exact", max_points=12, manual=dict(fit_range=150.0, fit_sill=3.0, fit_shape=2.4),
             use_nugget=True, nugget=0.2),
        dict(model="exponential", mode="estimate", precision=100, max_points=20),
        dict(model="spherical", mode="estimate", precision=800, max_points=15),
        dict(model="exponential", mode="estimate", precision=1000, max_points=32, use_nugget=True),
        dict(model="matern", mode="estimate", precision=200, max_points=16,
             manual=dict(fit_range=180.0, fit_sill=3.4, fit_shape=0.7)),
    ]
    for case in cases:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            extra = {}
            if "manual" in case:
                extra = dict(fit_method="manual", **case["manual"])
                if "nugget" in case:
                    extra["fit_nugget"] = case["nugget"]
            V = skg.Variogram(c, v, model=case["model"], n_lags=20, maxlag="median",
                              use_nugget=case.get("use_nugget", False), **extra)
            ok = skg.OrdinaryKriging(V, min_points=4, max_points=case["max_points"], mode=case["mode"],
                                     precision=case.get("precision", 100))
            z = ok.transform(t[:, 0], t[:, 1])
        d = V.describe()
        key = "e/%d" % len(manifest)
        arrays[key + "/z"] = z
        arrays[key + "/sigma"] = ok.sigma
        entry = dict(case)
        entry.update(key=key, effective_range=float(d["effective_range"]), sill=float(d["sill"]),
                     nugget=float(d["nugget"]),
                     shape=(float(d["smoothness"]) if "smoothness" in d else
                            (float(d["shape"]) if "shape" in d else None)),
                     no_points_error=int(ok.no_points_error), singular_error=int(ok.singular_error))
        manifest.append(entry)
        print(key, case["model"], case["mode"], "nan:", int(np.isnan(z).sum()), "of", len(z))
    # the Matern semivariance itself on a grid of (h, smoothness): pins the device Bessel function
    from skgstat import models as rm
    hs = np.concatenate(([0.0], np.geomspace(1e-3, 900.0, 120)))
    for k, sm in enumerate((0.2, 0.5, 0.7, 1.0, 1.5, 2.4, 3.0, 5.5, 10.0, 19.5)):
        arrays["matern/%d" % k] = np.array([rm.matern.py_func(h, 200.0, 3.0, sm, 0.1) if hasattr(rm.matern, "py_func")
                                            else rm.matern(np.array([h]), 200.0, 3.0, sm, 0.1)[0] for h in hs])
    arrays["matern/h"] = hs
    arrays["matern/s"] = np.array((0.2, 0.5, 0.7, 1.0, 1.5, 2.4, 3.0, 5.5, 10.0, 19.5))
    arrays["manifest"] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(OUT, "kriging_ext.npz"), **arrays)


def make_binners(skg, ds):
    """histogram-rule binners (binning.auto_derived_lags, binning.py:82-122)."""
    arrays, manifest = {}, []
    methods = ("sqrt", "sturges", "rice", "scott", "doane", "fd", "auto")
    for dname in ("pancake300", "rnd400", "lattice225", "rnd3d150", "rnd1d120", "gamma30"):
        c, v = ds[dname]
        arrays["in/%s/coords" % dname] = c
        arrays["in/%s/values" % dname] = v
        for method in methods:
            for ml in (None, "median", 0.6):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    V = skg.Variogram(c, v, bin_func=method, maxlag=ml, n_lags=10)
                if len(V.bins) > 240:
                    continue
                key = "b/%d" % len(manifest)
                arrays[key + "/bins"] = V.bins
                arrays[key + "/bin_count"] = np.asarray(V.bin_count, dtype=np.int64)
                arrays[key + "/experimental"] = V.experimental
                manifest.append(dict(key=key, kind="variogram", dataset=dname, bin_func=method,
                                     maxlag=ml, n_lags=int(V.n_lags)))
    c, v = ds["rnd400"]
    for method in ("sturges", "scott", "doane", "fd"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            V = skg.DirectionalVariogram(c, v, bin_func=method, maxlag="median", azimuth=30,
                                         tolerance=40.0, bandwidth=20.0)
        key = "b/%d" % len(manifest)
        arrays[key + "/bins"] = V.bins
        arrays[key + "/bin_count"] = np.asarray(V.bin_count, dtype=np.int64)
        arrays[key + "/experimental"] = V.experimental
        manifest.append(dict(key=key, kind="directional", dataset="rnd400", bin_func=method,
                             maxlag="median", n_lags=int(V.n_lags), azimuth=30, tolerance=40.0,
                             bandwidth=20.0))
    arrays["manifest"] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(OUT, "binners.npz"), **arrays)
    print("binner cases:", len(manifest))


def make_estimators(skg, ds):
    """minmax / percentile / entropy estimators (estimators.py:257-367,
    Variogram.py:2044-2092)."""
    arrays, manifest = {}, []
    est_cfgs = [("minmax", {}), ("percentile", {}), ("percentile", dict(percentile=25)),
                ("percentile", dict(percentile=90.5)), ("entropy", {}), ("entropy", dict(entropy_bins=20)),
                ("entropy", dict(entropy_bins="sturges"))]
    for dname in ("pancake300", "rnd400", "rnd3d150", "gamma30"):
        c, v = ds[dname]
        arrays["in/%s/coords" % dname] = c
        arrays["in/%s/values" % dname] = v
        for est, kw in est_cfgs:
            for bf, n_lags, ml in (("even", 10, None), ("uniform", 8, "median"), ("even", 15, "median")):
                try:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        V = skg.Variogram(c, v, estimator=est, bin_func=bf, n_lags=n_lags, maxlag=ml, **kw)
                        exp = V.experimental
                except Exception as e:      # e.g. minmax on an empty lag class raises in the reference
                    print("skip", dname, est, bf, ml, type(e).__name__)
                    continue
                key = "s/%d" % len(manifest)
                arrays[key + "/bins"] = V.bins
                arrays[key + "/bin_count"] = np.asarray(V.bin_count, dtype=np.int64)
                arrays[key + "/experimental"] = exp
                manifest.append(dict(key=key, kind="variogram", dataset=dname, estimator=est,
                                     bin_func=bf, n_lags=n_lags, maxlag=ml, kwargs=kw))
    c, v = ds["rnd400"]
    for est, kw in est_cfgs[:5]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            V = skg.DirectionalVariogram(c, v, estimator=est, n_lags=8, maxlag="median", azimuth=60,
                                         tolerance=45.0, bandwidth=25.0, **kw)
            exp = V.experimental
        key = "s/%d" % len(manifest)
        arrays[key + "/bins"] = V.bins
        arrays[key + "/bin_count"] = np.asarray(V.bin_count, dtype=np.int64)
        arrays[key + "/experimental"] = exp
        manifest.append(dict(key=key, kind="directional", dataset="rnd400", estimator=est, n_lags=8,
                             maxlag="median", azimuth=60, tolerance=45.0, bandwidth=25.0, kwargs=kw))
    arrays["manifest"] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(OUT, "estimators.npz"), **arrays)
    print("estimator cases:", len(manifest))


def make_sparse(skg, ds):
    """MetricSpace(max_dist=...) - only pairs with 0 < d <= max_dist exist
    (MetricSpace.py:141-147, Variogram.py:1202-1227, 1910-1936)."""
    arrays, manifest = {}, []
    sets = dict(ds)
    c, v = ds["rnd400"]
    cd = c.copy()
    cd[7] = cd[3]
    cd[100] = cd[3]          # coincident points: the sparse matrix drops their zero distances
    sets["rnd400dup"] = (cd, v)
    # the sparse branch of _format_values_stack subtracts in the values' own dtype (uint8 pancake
    # values wrap around there, unlike the dense pdist branch): use float observations
    sets["pancake300"] = (ds["pancake300"][0], np.asarray(ds["pancake300"][1], dtype=float))
    cfgs = [("rnd400", 30.0), ("rnd400", 61.5), ("rnd400dup", 45.0), ("pancake300", 150.0),
            ("lattice225", 5.0), ("rnd3d150", 22.0)]
    for dname, md in cfgs:
        c, v = sets[dname]
        arrays["in/%s/coords" % dname] = c
        arrays["in/%s/values" % dname] = v
        for est, bf, n_lags, ml in (("matheron", "even", 10, None), ("cressie", "uniform", 8, "median"),
                                    ("dowd", "even", 12, 0.5), ("matheron", "sturges", 10, None),
                                    ("matheron", "fd", 10, "median"), ("minmax", "even", 6, None),
                                    ("matheron", "even", 10, "mean"), ("matheron", "even", 8, md * 1.7)):
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ms = skg.MetricSpace(c, max_dist=md)
                    V = skg.Variogram(ms, v, estimator=est, bin_func=bf, n_lags=n_lags, maxlag=ml,
                                      fit_method=None)
                    key = "x/%d" % len(manifest)
                    out = dict(bins=V.bins, bin_count=np.asarray(V.bin_count, dtype=np.int64),
                               experimental=V.experimental,
                               maxlag=np.array(np.nan if V.maxlag is None else V.maxlag, dtype=float))
                    if est == "matheron" and bf == "even" and ml is None:
                        out["distance"] = V.distance
                        out["diffs"] = V.pairwise_diffs
            except Exception as e:   # e.g. minmax on an empty lag class raises in the reference
                print("skip", dname, md, est, bf, ml, type(e).__name__)
                continue
            for k2, v2 in out.items():
                arrays[key + "/" + k2] = v2
            manifest.append(dict(key=key, kind="variogram", dataset=dname, max_dist=md, estimator=est,
                                 bin_func=bf, n_lags=n_lags, maxlag=ml))
    # (DirectionalVariogram cannot take a MetricSpace in the reference: DirectionalVariogram.py:245
    #  reads self.dist_func before it exists)
    arrays["manifest"] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(OUT, "sparse.npz"), **arrays)
    print("sparse cases:", len(manifest))


def make_quality(skg):
    """Model quality measures of the reference (Variogram.py:2243-2553) on its own sample data
    (tests/sample.csv -> tests/golden/ref_sample.npz)."""
    z = np.load(os.path.join(OUT, "ref_sample.npz"))
    c, v = z["coords"], z["values"]
    out = {}
    for model in ("spherical", "exponential", "gaussian"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            V = skg.Variogram(c, v, model=model, n_lags=12, maxlag="median")
            out[model] = dict(mean_residual=float(V.mean_residual), root_mean_square=float(V.root_mean_square),
                              rss=float(V.rss), nrmse=float(V.nrmse), nrmse_r=float(V.nrmse_r), r=float(V.r),
                              NS=float(V.NS), rmse=float(V.rmse),
                              frame=[float(x) for x in V.to_DataFrame(n=7)[model].values])
    with open(os.path.join(OUT, "ref_quality.json"), "w") as fh:
        json.dump(out, fh, indent=1)


def make_api_fixture(skg):
    """The reference's own test fixture tests/sample.csv (200 rows x, y, z), used by
    test_variogram.py::test_fit_lm; data only."""
    import pandas as pd
    df = pd.read_csv(os.path.join("/root/reference/skgstat/tests", "sample.csv"))
    np.savez_compressed(os.path.join(OUT, "ref_sample.npz"),
                        coords=df[["x", "y"]].values, values=df.z.values)


def main():
    os.makedirs(OUT, exist_ok=True)
    skg = import_reference()
    ds = datasets(skg)
    which = sys.argv[1:] or ["variogram", "directional", "kriging", "cv", "api", "multivariate", "kriging_ext", "binners", "estimators", "sparse", "quality"]
    if "variogram" in which:
        make_variogram(skg, ds)
    if "directional" in which:
        make_directional(skg, ds)
    if "kriging" in which:
        make_kriging(skg, ds)
    if "cv" in which:
        make_cross_validation(skg, ds)
    if "api" in which:
        make_api_fixture(skg)
    if "multivariate" in which:
        make_multivariate(skg, ds)
    if "kriging_ext" in which:
        make_kriging_ext(skg, ds)
    if "binners" in which:
        make_binners(skg, ds)
    if "estimators" in which:
        make_estimators(skg, ds)
    if "sparse" in which:
        make_sparse(skg, ds)
    if "quality" in which:
        make_quality(skg)


if __name__ == "__main__":
    main()
