"""Seeded differential test: random small configurations (sizes, dimensions, duplicates, ties,
lag counts, maxlag forms, estimators, binners) through the drop-in classes against the dense
oracle.  Catches interactions the structured cases miss."""
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import variogram as ov

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def skg():
    import skgstat_b200
    skgstat_b200._lib.load()
    return skgstat_b200


def _config(rng):
    n = int(rng.choice([2, 3, 5, 17, 64, 65, 130, 255, 256, 257, 300, 511, 513, 700]))
    dim = int(rng.choice([1, 2, 2, 2, 3]))
    kind = rng.choice(["uniform", "lattice", "clustered", "dup"])
    if kind == "lattice":
        c = rng.integers(0, 12, size=(n, dim)).astype(float)
    elif kind == "clustered":
        c = rng.normal(0, 1, size=(n, dim)) + rng.integers(0, 3, size=(n, 1)) * 10.0
    else:
        c = rng.random((n, dim)) * float(rng.choice([1.0, 100.0, 1e4]))
        if kind == "dup" and n > 4:
            c[rng.integers(0, n, size=n // 4)] = c[0]
    v = rng.normal(size=n) * float(rng.choice([1.0, 50.0])) + float(rng.choice([0.0, 1e3]))
    if rng.random() < 0.2:
        v = np.round(v, 0)
    est = str(rng.choice(["matheron", "cressie", "dowd"]))
    bf = str(rng.choice(["even", "even", "uniform"]))
    n_lags = int(rng.choice([1, 2, 5, 10, 25, 40]))
    ml = rng.choice(["none", "median", "mean", "frac", "abs"])
    return c if dim > 1 else c[:, 0], v, est, bf, n_lags, str(ml)


def test_random_configurations_match_the_oracle(skg):
    rng = np.random.default_rng(20240607)
    done = 0
    for it in range(140):
        c, v, est, bf, n_lags, mlk = _config(rng)
        d = ov.distance(c)
        if not (np.nanmax(d) > 0):
            continue
        ml = {"none": None, "median": "median", "mean": "mean", "frac": 0.37,
              "abs": float(np.round(0.6 * d.max(), 6)) + 1.0}[mlk]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                ref = ov.dense_variogram(c, v, est, bf, n_lags, ml)
            except Exception:
                continue    # the reference itself cannot form these bins (e.g. empty percentile input)
            edges = np.asarray(ref["bins"], dtype=float)
            if np.any(np.diff(edges) < 0) or not np.all(np.isfinite(edges)):
                continue
            V = skg.Variogram(c, v, estimator=est, bin_func=bf, n_lags=n_lags, maxlag=ml, fit_method=None)
            msg = "iteration %d: n=%d est=%s bf=%s n_lags=%d maxlag=%r" % (it, len(v), est, bf, n_lags, ml)
            assert_array_equal(V.bins, ref["bins"], err_msg=msg)
            assert_array_equal(V.bin_count, ref["bin_count"], err_msg=msg)
            assert_allclose(V.experimental, ref["experimental"], rtol=1e-12, atol=1e-300, equal_nan=True,
                            err_msg=msg)
        done += 1
    assert done >= 100


def test_random_kriging_configurations_match_the_oracle(skg):
    """Random sample sets / targets / neighbourhood sizes through OrdinaryKriging against the
    oracle restatement of Kriging.py:405-650 (well-conditioned models only, rtol 1e-9)."""
    from oracle import kriging as okr
    rng = np.random.default_rng(777)
    done = 0
    for it in range(40):
        n = int(rng.choice([6, 20, 64, 65, 200, 500]))
        dim = int(rng.choice([2, 2, 3]))
        c = rng.random((n, dim)) * 100.0
        if rng.random() < 0.3 and n > 10:
            c[rng.integers(0, n, size=3)] = c[1]            # duplicated coordinates are removed
        v = np.sin(c[:, 0] / 15.0) + 0.1 * rng.normal(size=n)
        model = str(rng.choice(["spherical", "exponential", "cubic", "stable", "matern"]))
        rng_r = float(rng.choice([15.0, 40.0, 90.0]))
        sill = float(rng.choice([0.5, 2.0]))
        nug = float(rng.choice([0.0, 0.0, 0.05]))
        shape = {"stable": 1.2, "matern": 0.8}.get(model)
        minp = int(rng.choice([1, 3, 5]))
        maxp = int(rng.choice([5, 8, 16, 33, 64]))
        if minp > maxp:
            minp = maxp
        m = 60
        t = rng.random((m, dim)) * 120.0 - 10.0
        t[:5] = c[:5]                                       # targets on top of samples
        desc = dict(model=model, effective_range=rng_r, sill=sill, nugget=nug, dist_func="euclidean")
        if model == "stable":
            desc["shape"] = shape
        if model == "matern":
            desc["smoothness"] = shape
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ok = skg.OrdinaryKriging(desc, min_points=minp, max_points=maxp, coordinates=c, values=v)
            z = ok.transform(*[t[:, k] for k in range(dim)])
            zr, sr, st = okr.krige(c, v, t, model, rng_r, sill, nug, shape, minp, maxp)
        msg = "iteration %d: n=%d dim=%d model=%s range=%g maxp=%d" % (it, n, dim, model, rng_r, maxp)
        good = st == 0
        assert_array_equal(np.isnan(z), ~good | np.isnan(zr), err_msg=msg)
        fin = good & np.isfinite(zr)
        assert_allclose(z[fin], zr[fin], rtol=1e-8, atol=1e-9, err_msg=msg)
        assert_allclose(ok.sigma[fin], sr[fin], rtol=1e-7, atol=1e-8 * sill, err_msg=msg)
        done += 1
    assert done == 40


def test_random_directional_configurations_match_the_oracle(skg):
    """Random azimuths / tolerances / bandwidths (incl. the extremes) on random and lattice
    points: exercises the guard-band resolution and the direction culling of tiles and warps."""
    rng = np.random.default_rng(4242)
    done = 0
    for it in range(60):
        n = int(rng.choice([40, 257, 600, 1100, 2300]))
        if rng.random() < 0.3:
            c = rng.integers(0, 25, size=(n, 2)).astype(float)
        else:
            c = rng.random((n, 2)) * float(rng.choice([10.0, 1000.0]))
        v = rng.normal(size=n)
        az = float(rng.choice([0, 45, 90, 135, 180, -120, 30.5, -179.9, -45]))
        tol = float(rng.choice([1.0, 10.0, 22.5, 45.0, 90.0, 170.0, 180.0]))
        model = str(rng.choice(["triangle", "compass"]))
        span = float(c.max() - c.min())
        bw = rng.choice(["q33", "q10", "abs_small", "abs_large"])
        bw = {"q33": "q33", "q10": "q10", "abs_small": 0.02 * span, "abs_large": 2.0 * span}[str(bw)]
        est = str(rng.choice(["matheron", "cressie", "dowd"]))
        bf = str(rng.choice(["even", "uniform"]))
        n_lags = int(rng.choice([4, 10, 20]))
        ml = rng.choice(["none", "median", "frac"])
        ml = {"none": None, "median": "median", "frac": 0.5}[str(ml)]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                ref = ov.dense_directional(c, v, est, bf, n_lags, ml, az, tol, bw, model)
            except Exception:
                continue
            edges = np.asarray(ref["bins"], dtype=float)
            if ref["mask_count"] == 0 or np.any(np.diff(edges) < 0) or not np.all(np.isfinite(edges)):
                continue
            msg = ("iteration %d: n=%d az=%g tol=%g bw=%r model=%s est=%s bf=%s n_lags=%d maxlag=%r"
                   % (it, n, az, tol, bw, model, est, bf, n_lags, ml))
            try:
                V = skg.DirectionalVariogram(c, v, estimator=est, bin_func=bf, n_lags=n_lags, maxlag=ml,
                                             azimuth=az, tolerance=tol, bandwidth=bw,
                                             directional_model=model, fit_method=None)
                bins = V.bins
            except Exception as e:
                raise AssertionError(msg + " -> " + repr(e) + " (oracle mask_count=%d)" % ref["mask_count"])
            assert_array_equal(bins, ref["bins"], err_msg=msg)
            assert_array_equal(V.bin_count, ref["bin_count"], err_msg=msg)
            assert_allclose(V.experimental, ref["experimental"], rtol=1e-12, atol=1e-300, equal_nan=True,
                            err_msg=msg)
        done += 1
    assert done >= 40


def test_random_rule_binners_estimators_and_sparse_spaces_match_the_oracle(skg):
    """Random configurations over the histogram-rule binners, the minmax / percentile / entropy
    estimators and sparse (max_dist) spaces."""
    rng = np.random.default_rng(99)
    done = 0
    for it in range(90):
        n = int(rng.choice([30, 120, 256, 400, 777]))
        dim = int(rng.choice([2, 2, 3]))
        c = rng.random((n, dim)) * 100.0
        if rng.random() < 0.25:
            c[rng.integers(0, n, size=3)] = c[0]
        v = rng.normal(size=n) * 3.0 + 10.0
        est = str(rng.choice(["matheron", "cressie", "dowd", "minmax", "percentile", "entropy"]))
        bf = str(rng.choice(["even", "uniform", "sturges", "scott", "doane", "fd", "rice", "auto"]))
        n_lags = int(rng.choice([5, 10, 20]))
        ml = {0: None, 1: "median", 2: 0.45}[int(rng.integers(0, 3))]
        md = None if rng.random() < 0.5 else float(rng.choice([25.0, 60.0, 500.0]))
        kw = {}
        if est == "percentile" and rng.random() < 0.5:
            kw["percentile"] = float(rng.choice([10, 33.3, 75]))
        if est == "entropy" and rng.random() < 0.5:
            kw["entropy_bins"] = int(rng.choice([12, 30]))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                with np.errstate(all="ignore"):
                    ref = ov.dense_variogram(c, v, est, bf, n_lags, ml, max_dist=md, **kw)
            except Exception:
                continue    # the reference raises here (e.g. minmax / percentile of an empty lag class)
            edges = np.asarray(ref["bins"], dtype=float)
            if edges.size == 0 or edges.size > 240 or np.any(np.diff(edges) < 0) or not np.all(np.isfinite(edges)):
                continue
            coords = c if md is None else skg.MetricSpace(c, max_dist=md)
            V = skg.Variogram(coords, v, estimator=est, bin_func=bf, n_lags=n_lags, maxlag=ml, fit_method=None, **kw)
            msg = "iteration %d: n=%d dim=%d est=%s bf=%s n_lags=%d maxlag=%r max_dist=%r %r" % (
                it, n, dim, est, bf, n_lags, ml, md, kw)
            assert_array_equal(V.bins, ref["bins"], err_msg=msg)
            assert_array_equal(V.bin_count, ref["bin_count"], err_msg=msg)
            assert_allclose(V.experimental, ref["experimental"], rtol=1e-12, atol=1e-300, equal_nan=True,
                            err_msg=msg)
        done += 1
    assert done >= 50
