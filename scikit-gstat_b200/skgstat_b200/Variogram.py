"""Variogram: the reference's experimental-variogram API on top of the fused
B200 pair kernels.

Same constructor and properties as ``skgstat.Variogram`` (Variogram.py:31-48)
for the path BASELINE.json names: ``Variogram(coordinates, values, estimator,
model, dist_func, bin_func, normalize, fit_method, fit_sigma, use_nugget,
maxlag, samples, n_lags, ...)`` with ``bins``, ``bin_count``, ``experimental``,
``maxlag``, ``n_lags``, ``fit()``, ``describe()``, ``parameters`` and the cache
invalidation contract of its setters.

What differs underneath: nothing O(N^2) is ever stored.  ``maxlag='median'``,
ratio maxlags and ``bin_func='uniform'`` get exact order statistics / the exact
maximum from select sweeps (engine.sq_order_stats), lag edges are then formed
with numpy's own arithmetic from those scalars (bit-identical edges), and
``experimental`` / ``bin_count`` come from ONE fused sweep
(engine.hist -> skg_pair_hist) or, for Dowd, from the per-lag median select
(engine.dz_medians).  The materialising accessors (``distance``,
``pairwise_diffs``, ``lag_groups()``, ``lag_classes()``) remain, produced on the
device on demand.  The model fit is scipy ``curve_fit`` on n_lags points and
stays on the CPU, as in the reference (Variogram.py:1584-1821).

Also built on the same kernels: cross / aggregate (2-D ``values``) variograms, the
histogram-rule binners, the minmax / percentile / entropy estimators, sparse ``max_dist`` spaces.

Out of scope, raising NotImplementedError rather than taking a slow path:
``samples=`` (ProbabalisticMetricSpace), non-euclidean metrics, the clustering binners
(kmeans, ward, stable_entropy), the genton estimator by name and harmonised / sum models
(SURVEY.md section 2 rows 8-15).
"""
from __future__ import annotations

import copy
import warnings
from collections.abc import Callable, Iterable

import numpy as np
from scipy.optimize import curve_fit

from . import _lib, binning, estimators, models
from ._exact import sq_threshold_bits
from .MetricSpace import MetricSpace

_GPU_ESTIMATORS = ("matheron", "cressie", "dowd", "minmax", "percentile", "entropy")
_GPU_BINNERS = ("even", "uniform")


class Variogram:
    _token_counter = 0

    def __init__(self, coordinates=None, values=None, estimator="matheron", model="spherical",
                 dist_func="euclidean", bin_func="even", normalize=False, fit_method="trf",
                 fit_sigma=None, use_nugget=False, maxlag=None, samples=None, n_lags=10,
                 multivariate="cross", verbose=False, **kwargs):
        self._process_group = kwargs.pop("process_group", None)
        self._kwargs = dict(kwargs)
        if samples is not None:
            raise NotImplementedError(
                "samples= (ProbabalisticMetricSpace) is out of scope: the dense sweep it "
                "approximates is what the GPU path accelerates")
        self._1d = False
        if not isinstance(coordinates, MetricSpace):
            c = np.asarray(coordinates)
            if c.ndim < 2:  # Variogram.py:291-297
                c = np.column_stack((c, np.zeros(len(c))))
                self._1d = True
            coordinates = MetricSpace(c.copy(), dist_func, None)
        elif dist_func != coordinates.dist_metric:
            raise AttributeError("Distance metric of variogram differs from distance metric "
                                 "of coordinates")
        self._X = coordinates
        self.multivariate = multivariate
        self.verbose = verbose
        self._reset_pair_state()
        # set_values decides whether this is a cross / aggregate variogram (Variogram.py:328-332)
        self._is_cross = None
        self._co_variable = None
        self._is_aggregate = None
        self._values = None
        self.set_values(values)

        self._n_lags_passed_value = n_lags
        self._n_lags = None
        self.n_lags = n_lags
        self._maxlag = None
        self._maxlag_within_range = False
        self.maxlag = maxlag

        self._estimator = None
        self._estimator_name = None
        self.set_estimator(estimator)
        self._bin_func_name = None
        self._bin_func = None
        self._bins = None
        self.set_bin_func(bin_func)

        self._use_nugget = None
        self.use_nugget = use_nugget
        self._model = None
        self._model_name = None
        self._is_model_custom = False
        self.set_model(model)
        self._normalized = normalize
        self._fit_method = None
        if fit_method is not None:
            self.fit_method = fit_method
        self._fit_sigma = None
        self.fit_sigma = fit_sigma
        self.cov = None
        self.cof = None
        self.fit(force=True, bounds=self._kwargs.get("fit_bounds"), p0=self._kwargs.get("fit_p0"))

    # ------------------------------------------------------------------ state
    def _reset_pair_state(self):
        """Drop everything derived from (bins, values, mask): the reference resets
        _groups/_bin_count/cof/cov in every setter (Variogram.py:885-940, 1274-1282)."""
        self._groups = None
        self._bin_count = None
        self._exp = None
        self.cof = None
        self.cov = None

    def _dirp(self):
        """Directional mask parameters for the kernels (None = isotropic)."""
        return None

    def _dir_key(self):
        return None

    def _engine(self):
        eng = self._X.engine(self._process_group)
        # token, not a reference to self: no reference cycle, so device buffers are
        # released as soon as the variogram goes away
        if getattr(eng, "_values_owner", None) != self._values_token or eng.values is None:
            eng.set_values(*self._engine_values())
            eng._values_owner = self._values_token
        return eng

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_process_group"] = None
        return st

    # ------------------------------------------------------------- properties
    @property
    def coordinates(self):
        return self._X.coords

    @property
    def metric_space(self):
        return self._X

    @property
    def dim(self):
        return 1 if self._1d else self.coordinates.shape[1]

    @property
    def values(self):
        return self._values

    @values.setter
    def values(self, values):
        self.set_values(values)

    def set_values(self, values, calc_diff=True):
        """Variogram.py:499-590: 1-D observations, (N, 2) observations + co-variable
        (cross-variogram, |dz_0|*|dz_1|) or (N, k) with multivariate='aggregate'."""
        if not len(values) == len(self.coordinates):
            raise ValueError("The length of the values array has to match the length of "
                             "coordinates")
        y = np.asarray(values)
        if y.ndim > 2:
            raise ValueError("values has to be 1d (classic variogram) or 2d dimensional "
                             "(cross-variogram)")
        if y.ndim == 2:
            if self.multivariate == "aggregate":
                warnings.warn("Perform analysis on %d dimension data" % y.ndim)
                if y.shape[1] > _lib.MAX_VALUE_COLUMNS:
                    raise NotImplementedError("multivariate='aggregate' supports up to %d columns "
                                              "on the GPU path" % _lib.MAX_VALUE_COLUMNS)
                self._is_cross = False
                self._is_aggregate = True
            elif self.multivariate == "cross":
                if y.shape[1] > 2:
                    raise ValueError("Use the utility function to create a grid of cross-variograms")
                elif y.shape[1] == 2:
                    self._co_variable = y[:, 1].flatten()
                    if self._is_cross is not None and not self._is_cross:
                        warnings.warn("You passed two variables as observation and effectively "
                                      "turned this instance into a cross-variogram.")
                    self._is_cross = True
                    self._is_aggregate = False
                    y = y[:, 0].flatten()
                # (N, 1) falls through unchanged, as in the reference
            else:
                raise NotImplementedError("Multivariate mode %s is not implemented for this case."
                                          % self.multivariate)
        else:
            self._is_cross = False
            self._is_aggregate = False
        flat = y.ravel()
        self._single_input = flat.size == 0 or bool(np.all(flat == flat[0]))
        if self._single_input:  # len(set(values)) < 2, Variogram.py:575
            warnings.warn("All input values are the same.")
        self._values = y
        self._values_token = (id(self), Variogram._token_counter)
        Variogram._token_counter += 1
        self._diff = None
        self._reset_pair_state()

    def _engine_values(self):
        """(array, value_mode) handed to the pair engine."""
        if self._is_cross:
            return np.column_stack((self._values, self._co_variable)), _lib.VALUE_CROSS
        if self._is_aggregate:
            return self._values, _lib.VALUE_AGGREGATE
        v = self._values
        return (v[:, 0] if v.ndim == 2 else v), _lib.VALUE_SCALAR

    @property
    def is_cross_variogram(self):
        """Read-only flag: cross-variogram (Variogram.py:1468-1471)."""
        return self._is_cross

    @property
    def is_agg_variogram(self):
        """Read-only flag: aggregated multi-column data (Variogram.py:1473-1476)."""
        return self._is_aggregate

    @property
    def n_lags(self):
        if self._n_lags is None:
            self._n_lags = len(self.bins)
        return self._n_lags

    @n_lags.setter
    def n_lags(self, n):
        if isinstance(n, str):
            raise NotImplementedError("n_lags string values not implemented")
        if isinstance(n, bool) or not isinstance(n, (int, np.integer)) or n < 1:
            raise ValueError("n_lags has to be a positive integer")
        if n > 100000:
            raise ValueError("n_lags > 100000 is not supported")
        self._n_lags = int(n)
        self._n_lags_passed_value = n
        self._bins = None
        self._reset_pair_state()

    @property
    def maxlag(self):
        return self._maxlag

    @maxlag.setter
    def maxlag(self, value):
        """Variogram.py:1274-1295."""
        self._bins = None
        self._reset_pair_state()
        self._maxlag_within_range = False
        if value is None:
            self._maxlag = None
        elif isinstance(value, str):
            if value == "median":
                self._maxlag = self._distance_median()
                self._maxlag_within_range = True
            elif value == "mean":
                # np.mean sums the condensed vector pairwise in its storage order; the
                # only way to get that exact value is to form the vector (as the reference)
                self._maxlag = np.mean(self.distance)
        elif value < 1:
            self._maxlag = value * self._distance_max()
            self._maxlag_within_range = value >= 0
        else:
            self._maxlag = value

    # ---- exact scalars of the distance distribution (never materialised) -----
    def _distance_max(self, masked=False):
        """np.max / np.nanmax of the (mask-passing) distances."""
        dirp = self._dirp() if masked else None
        key = ("dmax", self._dir_key() if masked else None)

        def compute():
            s = self._X.engine(self._process_group).sq_max(dirp)
            return None if s is None else float(np.sqrt(np.float64(s)))

        return self._X.cached(key, compute)

    def _distance_order_stats(self, rank_fn, masked=False, limit=None):
        dirp = self._dirp() if masked else None
        eng = self._X.engine(self._process_group)
        vals, total = eng.sq_order_stats(rank_fn, dirp=dirp, limit=limit)
        return {r: float(np.sqrt(np.float64(s))) for r, s in vals.items()}, total

    def _distance_median(self):
        def compute():
            d, total = self._distance_order_stats(
                lambda t: [] if t == 0 else [(t - 1) // 2, t // 2])
            if total == 0:
                return np.nan
            return binning.median_from_order_stats(total, lambda r: d[r])
        return self._X.cached(("dmedian",), compute)

    def _distance_percentile(self, q):
        def compute():
            d, total = self._distance_order_stats(
                lambda t: [] if t == 0 else list(binning.percentile_ranks(t, q)[:2]))
            if total == 0:
                return np.nan
            return binning.percentile_from_order_stats(total, q, lambda r: d[r])
        return self._X.cached(("dpercentile", float(q)), compute)

    # ------------------------------------------------------------------ binning
    @property
    def bin_func(self):
        return self._bin_func

    @bin_func.setter
    def bin_func(self, bin_func):
        self.set_bin_func(bin_func)

    def set_bin_func(self, bin_func):
        """Variogram.py:771-817."""
        if isinstance(bin_func, str):
            name = bin_func.lower()
            if name in binning.AUTO_BINNERS:
                # np.histogram_bin_edges rules (binning.py:82-122): the rule's inputs (count, min,
                # max, quartiles, moments of the distances) come from the pair kernels
                self._bin_func = name
                self._bin_func_name = bin_func
                self._bins = None
                self._reset_pair_state()
                return
            if name not in _GPU_BINNERS:
                if name in ("stone", "kmeans", "ward", "stable_entropy"):
                    raise NotImplementedError(
                        "bin_func=%r needs every distance in host memory and is out of "
                        "scope of the GPU path (even / uniform / callable / edges)" % bin_func)
                raise ValueError("%r is not a valid estimator for `bins`" % bin_func)
            self._bin_func = name
            self._bin_func_name = bin_func
            self._bins = None
        elif isinstance(bin_func, Callable):
            self._bin_func = bin_func
            self._bin_func_name = "custom_func"
            self._bins = None
        elif isinstance(bin_func, Iterable):
            edges = np.asarray(list(bin_func), dtype=float)
            self._bin_func = None
            self._bin_func_name = "custom_bin_edges"
            self._bins = edges
            self._maxlag = max(edges)
            self._n_lags = len(edges)
        else:
            raise AttributeError("bin_func has to be of type string, iterable or callable.")
        self._reset_pair_state()

    def _calc_bins(self):
        """Upper lag-class edges (binning.py:10-79; DirectionalVariogram.py:577-591)."""
        n, maxlag = self._n_lags, self._maxlag
        masked = self._dirp() is not None
        if self._bin_func == "even":
            if maxlag is not None and not masked and self._maxlag_within_range:
                # maxlag came from the distance distribution itself (median / ratio of the
                # max): the clamp `maxlag > nanmax(d)` of binning.py:37 cannot trigger
                return np.linspace(0, maxlag, n + 1)[1:], None
            dmax = self._distance_max(masked)
            if dmax is None:
                raise ValueError("no point pair passes the directional mask")
            return binning.even_edges(n, maxlag, dmax), None
        if self._bin_func == "uniform":
            qs = [(i / n) * 100 for i in range(1, n + 1)]

            def rank_fn(total):
                if total == 0:
                    return []
                out = set()
                for q in qs:
                    out.update(binning.percentile_ranks(total, q)[:2])
                return sorted(out)

            d, total = self._distance_order_stats(rank_fn, masked=masked, limit=maxlag)
            if total == 0:
                raise ValueError("no point pair inside maxlag / the directional mask")
            return np.fromiter((binning.percentile_from_order_stats(total, q, lambda r: d[r])
                                for q in qs), dtype=float), None
        if self._bin_func in binning.AUTO_BINNERS:
            return self._auto_bins(self._bin_func, maxlag, masked)
        # user callable: the reference's operator contract hands it the materialised vector
        d = self.distance.copy()
        if masked:
            d[np.where(~self._direction_mask())] = np.nan
        return self._bin_func(d, n, maxlag)

    def _auto_bins(self, method, maxlag, masked, limit_check=True):
        """binning.auto_derived_lags (binning.py:82-122) without the distance vector: exact
        count / min / max / quartiles from the select passes, std and skewness from three
        moment sweeps; the edges are np.linspace(min, max, n_bins + 1)[1:] as in numpy."""
        dmax = self._distance_max(masked)
        if dmax is None:
            raise ValueError("no point pair passes the directional mask")
        if maxlag is None or maxlag > dmax:
            maxlag = dmax
        need_q = method in ("fd", "auto")

        def rank_fn(total):
            if total == 0:
                return []
            out = {0, total - 1}
            if need_q:
                for q in (25, 75):
                    out.update(binning.percentile_ranks(total, q)[:2])
            return sorted(out)

        d, total = self._distance_order_stats(rank_fn, masked=masked, limit=maxlag)
        if total == 0:
            raise ValueError("no point pair inside maxlag / the directional mask")
        stats = {}
        if need_q:
            stats["q25"] = binning.percentile_from_order_stats(total, 25, lambda r: d[r])
            stats["q75"] = binning.percentile_from_order_stats(total, 75, lambda r: d[r])
        if method in ("scott", "doane"):
            eng = self._X.engine(self._process_group)
            n, s1, s2, s3 = eng.dist_moments(maxlag, self._dirp() if masked else None,
                                             third=(method == "doane"))
            if n != total:
                raise RuntimeError("moment sweep and select pass disagree on the pair count")
            mean = s1 / n
            var = max(s2 / n - mean * mean, 0.0)
            stats["std"] = float(np.sqrt(var))
            if method == "doane":
                m3 = s3 / n - 3.0 * mean * (s2 / n) + 2.0 * mean ** 3
                stats["g1"] = m3 / stats["std"] ** 3 if stats["std"] > 0 else 0.0
        edges = binning.auto_edges(method, total, d[0], d[total - 1], **stats)
        if limit_check and len(edges) > 100000:
            raise ValueError("bin_func=%r derives %d lag classes; at most 100000 are supported"
                             % (method, len(edges)))
        return edges, len(edges)

    @property
    def bins(self):
        if self._bins is None:
            edges, n = self._calc_bins()
            self._bins = np.asarray(edges, dtype=float)
            if n is not None:
                self._n_lags = n
        return self._bins.copy()

    @bins.setter
    def bins(self, bins):
        self._bins = np.asarray(bins, dtype=float)
        self._reset_pair_state()

    # ---------------------------------------------------------------- estimator
    @property
    def estimator(self):
        return self._estimator

    @estimator.setter
    def estimator(self, value):
        self.set_estimator(value)

    def set_estimator(self, estimator_name):
        """Variogram.py:957-986."""
        self._exp = None
        self.cof, self.cov = None, None
        if isinstance(estimator_name, str):
            name = estimator_name.lower()
            if name in _GPU_ESTIMATORS:
                self._estimator = estimators.ESTIMATORS[name]
                self._estimator_name = name
            elif name == "genton":
                raise NotImplementedError(
                    "estimator 'genton' needs all pairs of pair differences per lag class and is "
                    "outside the accelerated path; pass estimators.genton-like callables to run "
                    "them on materialised lag classes")
            else:
                raise ValueError("Variogram estimator %s is not understood, please provide "
                                 "the function." % estimator_name)
        elif callable(estimator_name):
            self._estimator = estimator_name
            self._estimator_name = None
        else:
            raise ValueError("The estimator has to be a string or callable.")

    LAG_CHUNK = 240   # lag classes per sweep (the kernels' lag table holds SKG_MAX_LAGS = 250)

    def _calc_experimental(self):
        if self._exp is not None:
            return
        edges = self.bins
        thr = sq_threshold_bits(edges)
        if np.any(thr[1:] < thr[:-1]):
            raise ValueError("bins must be monotonically increasing")
        eng = self._engine()
        dirp = self._dirp()
        if self._estimator_name is None:
            # user callable on materialised lag classes (reference operator contract)
            classes = list(self.lag_classes())
            cnt = np.fromiter((c.size for c in classes), dtype=np.int64)
            exp = np.fromiter(map(self._estimator, classes), dtype=float)
        elif len(edges) <= self.LAG_CHUNK:
            cnt, exp = self._experimental_of(eng, edges, thr, dirp)
        else:
            # more lag classes than one sweep holds: sweeps over consecutive groups of classes, each
            # with the previous edge as a leading class whose result is dropped (tiles beyond a
            # group's last edge are culled, so k groups cost about k/2 full sweeps)
            cnts, exps = [], []
            step = self.LAG_CHUNK - 1
            for a in range(0, len(edges), step):
                z = min(len(edges), a + step)
                lead = 1 if a else 0
                c, e = self._experimental_of(eng, edges[a - lead:z], thr[a - lead:z], dirp)
                cnts.append(c[lead:])
                exps.append(e[lead:])
            cnt, exp = np.concatenate(cnts), np.concatenate(exps)
        self._bin_count = np.asarray(cnt, dtype=np.int64)
        self._exp = np.asarray(exp, dtype=float)

    def _experimental_of(self, eng, edges, thr, dirp):
        """(pair counts, semi-variances) of the lag classes with the given upper edges."""
        name = self._estimator_name
        if name == "matheron":
            cnt, ssq, _ = eng.hist(edges, _lib.STAT_SUM_SQ, dirp, thr_bits=thr)
            exp = estimators.finalize_matheron(cnt, ssq)
        elif name == "cressie":
            cnt, _, srt = eng.hist(edges, _lib.STAT_SUM_SQRT, dirp, thr_bits=thr)
            exp = estimators.finalize_cressie(cnt, srt)
        elif name == "dowd":
            cnt, med = eng.dz_medians(edges, dirp, thr_bits=thr)
            exp = estimators.finalize_dowd(cnt, med)
        elif name == "minmax":  # estimators.py:296: (nanmax - nanmin) / nanmean per lag class
            cnt, total = eng.dz_abs_sums(edges, dirp, thr_bits=thr)
            vals, tot = eng.dz_order_stats(edges, lambda g, t: [0, t - 1] if t else [], dirp, thr_bits=thr)
            lo = [vals[g][0] if tot[g] else np.nan for g in range(len(edges))]
            hi = [vals[g][tot[g] - 1] if tot[g] else np.nan for g in range(len(edges))]
            exp = estimators.finalize_minmax(cnt, lo, hi, total)
        elif name == "percentile":  # Variogram.py:2076-2082: kwargs['percentile'] or 50
            p = self._kwargs.get("percentile") if self._kwargs.get("percentile", False) else 50
            vals, tot = eng.dz_order_stats(
                edges, lambda g, t: sorted(set(binning.percentile_ranks(t, p)[:2])) if t else [],
                dirp, thr_bits=thr)
            cnt = np.asarray(tot, dtype=np.int64)
            exp = np.array([binning.percentile_from_order_stats(tot[g], p, lambda r, g=g: vals[g][r])
                            if tot[g] else np.nan for g in range(len(edges))])
        else:  # "entropy": Variogram.py:2065-2075 + util/shannon.py:27-34
            value_edges = self._entropy_value_edges()
            hist = eng.dz_value_histogram(edges, value_edges, dirp, thr_bits=thr)
            cnt = eng.hist_counts(thr, dirp)
            exp = np.array([estimators.shannon_from_counts(hist[g]) for g in range(len(edges))])
        if name in ("dowd", "minmax", "percentile") and not np.isfinite(np.asarray(self._values, dtype=float)).all():
            # the order-statistic passes count finite differences only; a lag class's size
            # (bin_count, Variogram.py:942-947) includes the pairs of NaN observations
            cnt = eng.hist_counts(thr, dirp)
        return np.asarray(cnt, dtype=np.int64), np.asarray(exp, dtype=float)

    def experimental_batch(self, values):
        """Experimental variograms of K value vectors observed at THIS instance's coordinates,
        with its bins, estimator and directional mask: values (K, N) -> (K, n_lags).

        matheron / cressie: four vectors share one pair sweep (engine.hist_batch), which is how
        util.uncertainty.propagate and the diagonal of util.cross_variogram.cross_variograms
        avoid re-running the distance / lag-class work per vector.  Other estimators fall back
        to one pass per vector."""
        V = np.asarray(values, dtype=float)
        if V.ndim != 2 or V.shape[1] != len(self.coordinates):
            raise ValueError("values must have shape (K, %d)" % len(self.coordinates))
        edges = self.bins
        thr = sq_threshold_bits(edges)
        eng = self._X.engine(self._process_group)
        name = self._estimator_name
        if name in ("matheron", "cressie") and len(edges) <= min(eng.lib.skg_pair_batch_max_lags(),
                                                                  self.LAG_CHUNK - 2):
            stat = _lib.STAT_SUM_SQ if name == "matheron" else _lib.STAT_SUM_SQRT
            cnt, sums = eng.hist_batch(edges, V, stat, self._dirp(), thr_bits=thr)
            fin = estimators.finalize_matheron if name == "matheron" else estimators.finalize_cressie
            return np.vstack([fin(cnt, sums[k]) for k in range(V.shape[0])])
        out = np.empty((V.shape[0], len(edges)))
        keep = (self._values, self._co_variable, self._is_cross, self._is_aggregate, self._exp,
                self._bin_count, self.cof, self.cov, self._values_token, self._diff, self._groups)
        try:
            for k in range(V.shape[0]):
                self._is_cross, self._is_aggregate, self._co_variable = False, False, None
                self._values = V[k]
                self._values_token = (id(self), Variogram._token_counter)
                Variogram._token_counter += 1
                self._exp = None
                self._diff = None
                self._calc_experimental()
                out[k] = self._exp
        finally:
            (self._values, self._co_variable, self._is_cross, self._is_aggregate, self._exp,
             self._bin_count, self.cof, self.cov, self._values_token, self._diff, self._groups) = keep
        return out

    def _entropy_value_edges(self):
        """Variogram.py:2066-2073: np.histogram_bin_edges(self.distance, bins=entropy_bins - 1) -
        the reference derives the |dz| histogram edges from the DISTANCE vector; built here from
        the exact min / max (and, for rule names, quartiles / moments) of the distances."""
        nb = self._kwargs.get("entropy_bins", 50)
        d, total = self._distance_order_stats(lambda t: [0, t - 1] if t else [])
        if total == 0:
            raise ValueError("no point pairs")
        dmin, dmax = d[0], d[total - 1]
        if isinstance(nb, (int, np.integer)) and not isinstance(nb, bool):
            nb = int(nb) - 1
            if nb < 1:
                raise ValueError("`bins` must be positive, when an integer")
            first, last = np.float64(dmin), np.float64(dmax)
            if first == last:
                first, last = first - 0.5, last + 0.5
            return np.linspace(first, last, nb + 1, endpoint=True, dtype=np.float64)
        if isinstance(nb, str):
            edges, _ = self._auto_bins(nb.lower(), None, False, limit_check=False)
            return np.concatenate(([np.float64(dmin) if dmin != dmax else dmin - 0.5], edges))
        return np.asarray(nb, dtype=float)

    @property
    def bin_count(self):
        self._calc_experimental()
        return self._bin_count

    @property
    def experimental(self):
        self._calc_experimental()
        return self._exp.copy()

    @property
    def _experimental(self):
        return self.experimental

    def get_empirical(self, bin_center=False):
        edges = self.bins
        if bin_center:
            widths = np.diff(np.concatenate(([0.0], edges)))
            edges = edges - widths / 2
        return edges, self.experimental

    # --------------------------------------------------- materialising accessors
    @property
    def distance(self):
        """Condensed distance vector (Variogram.py:1202-1208); device-generated."""
        return self._X.engine(self._process_group).materialize(("dist",))["dist"]

    @property
    def distance_matrix(self):
        return self._X.dists

    @property
    def pairwise_diffs(self):
        if self._diff is None:
            self._diff = self._engine().materialize(("diffs",))["diffs"]
        return self._diff

    @property
    def value_matrix(self):
        from scipy.spatial.distance import squareform  # layout helper only
        return squareform(self.pairwise_diffs, checks=False)

    def _direction_mask(self):
        return np.ones(self._X.engine(self._process_group).n_pairs, dtype=bool)

    def lag_groups(self):
        """Lag class of every pair, -1 outside (Variogram.py:1989-2006)."""
        if self._groups is None and len(self.bins) > self.LAG_CHUNK:
            # materialising accessor beyond the kernels' lag table: digitize the materialised vector
            g = np.digitize(self.distance, self.bins)
            g = np.where(g == len(self.bins), -1, g)
            if self._dirp() is not None:
                g[~self._direction_mask()] = -1
            self._groups = g
        if self._groups is None:
            out = self._engine().materialize(("groups",), edges=self.bins, dirp=self._dirp())
            self._groups = out["groups"]
        return self._groups

    def lag_classes(self):
        diffs = self.pairwise_diffs
        groups = self.lag_groups()
        for i in range(len(self.bins)):
            yield diffs[np.where(groups == i)]

    def preprocessing(self, force=False):
        """Variogram.py:1559-1582: (re)compute bins and per-lag statistics."""
        if force:
            self._exp = None
            self._bin_count = None
            self._groups = None
        self._calc_experimental()

    # -------------------------------------------------------------------- model
    @property
    def model(self):
        return self._model

    @model.setter
    def model(self, value):
        self.set_model(value)

    def set_model(self, model_name):
        """Variogram.py:996-1030 (single named model or a callable)."""
        self.cof, self.cov = None, None
        if isinstance(model_name, str):
            name = model_name.lower()
            if "+" in name or name == "harmonize":
                raise NotImplementedError("sums of models / harmonize are CPU-side features "
                                          "outside the accelerated path")
            if name not in models.MODELS:
                raise ValueError("The theoretical Variogram function %s is not understood, "
                                 "please provide the function" % model_name)
            self._model = models.MODELS[name]
            self._model_name = name
            self._is_model_custom = False
        elif callable(model_name):
            self._model = model_name
            self._model_name = getattr(model_name, "__name__", "custom")
            self._is_model_custom = True
        else:
            raise ValueError("model has to be a string or a callable")

    @property
    def use_nugget(self):
        return self._use_nugget

    @use_nugget.setter
    def use_nugget(self, enable_nugget):
        if not isinstance(enable_nugget, bool):
            raise ValueError("use_nugget has to be of type bool.")
        self._use_nugget = enable_nugget

    @property
    def normalized(self):
        return self._normalized

    @normalized.setter
    def normalized(self, status):
        self._normalized = status

    @property
    def dist_function(self):
        return self._X.dist_metric

    @property
    def fit_method(self):
        return self._fit_method

    @fit_method.setter
    def fit_method(self, value):
        """Variogram.py:1334-1353."""
        if value not in ("lm", "ml", "trf", "manual"):
            raise AttributeError("fit_method has to be one of ['lm', 'ml', 'trf', 'manual']")
        if value == "trf" and self._single_input:
            raise AttributeError("'trf' is bounded and therefore not supported when all values "
                                 "are the same.")
        self._fit_method = value

    @property
    def fit_sigma(self):
        """Variogram.py:1359-1458 (distance-derived weights)."""
        s = self._fit_sigma
        if s is None:
            return None
        if isinstance(s, (list, tuple, np.ndarray)):
            if len(s) == len(self.bins):
                return s
            raise AttributeError("len(fit_sigma) != len(bins)")
        w = self.bins / np.max(self.bins)
        if s == "linear":
            return w
        if s == "exp":
            return 1.0 / np.exp(1.0 / w)
        if s == "sqrt":
            return np.sqrt(w)
        if s == "sq":
            return w ** 2
        raise ValueError("fit_sigma is not understood. It has to be an array or one of "
                         "['linear', 'exp', 'sqrt', 'sq'].")

    @fit_sigma.setter
    def fit_sigma(self, sigma):
        self._fit_sigma = sigma
        self.cof, self.cov = None, None

    def _fit_bounds(self, x, y):
        """Upper bounds: range <= max(bins), sill <= max(experimental), shape limits
        (Variogram.py:2137-2185)."""
        b = [np.nanmax(x), np.nanmax(y)]
        if self._model_name == "matern":
            b.append(20.0)
        elif self._model_name == "stable":
            b.append(2.0)
        if self.use_nugget:
            b.append(0.99 * np.nanmax(y))
        return b

    def fit(self, force=False, method=None, sigma=None, bounds=None, p0=None, **kwargs):
        """Least-squares fit of the theoretical model to (bins, experimental) on the CPU
        (Variogram.py:1584-1821): n_lags points, scipy curve_fit."""
        old = {} if self.cof is None else self.describe()
        if force:
            self.cof, self.cov = None, None
        self.preprocessing(force=force)
        if method is not None:
            self.fit_method = method
        if sigma is not None:
            self.fit_sigma = sigma
        if self._fit_method is None:
            return
        x = np.array(self.bins)
        y = self.experimental
        keep = ~np.isnan(y)
        _x, _y = x[keep], y[keep]
        if self._fit_method == "manual":
            if kwargs.get("nugget", False):
                self.use_nugget = True
            r = kwargs.get("range", self._kwargs.get("fit_range", old.get("effective_range")))
            s = kwargs.get("sill", self._kwargs.get("fit_sill", old.get("sill")))
            if r is None or s is None:
                raise AttributeError("For manual fitting, you need to pass the variogram "
                                     "parameters either to fit or to the Variogram instance.\n"
                                     "parameter need to be prefixed with fit_ if passed to "
                                     "__init__.")
            n = kwargs.get("nugget", self._kwargs.get("fit_nugget", old.get("nugget", 0.0)))
            if self._model_name in ("stable", "matern"):
                key = "shape" if self._model_name == "stable" else "smoothness"
                s2 = kwargs.get("shape", self._kwargs.get("fit_shape", old.get(key, 2.0)))
                self.cof = [r, s, s2, n]
            else:
                self.cof = [r, s, n]
            return
        if self._is_model_custom:
            def wrapped(h, *args):
                return self._model(h, *args)
            if bounds is None or p0 is None:
                raise ValueError("a custom model needs explicit fit bounds and p0")
        else:
            if self.use_nugget:
                def wrapped(h, *args):
                    return self._model(h, *args)
            else:
                def wrapped(h, *args):
                    return self._model(h, *args, 0)
            if bounds is None:
                bounds = (0, self._fit_bounds(x, y))
            if p0 is None:
                p0 = np.asarray(bounds[1])
        sig = self.fit_sigma
        if sig is not None:
            sig = np.asarray(sig)[keep]
        if self._fit_method == "trf":
            self.cof, self.cov = curve_fit(wrapped, _x, _y, method="trf", sigma=sig, p0=p0,
                                           bounds=bounds, **kwargs)
        elif self._fit_method == "lm":
            self.cof, self.cov = curve_fit(wrapped, _x, _y, method="lm", sigma=sig, p0=p0,
                                           **kwargs)
        else:  # 'ml': maximise the gaussian log-likelihood of the lag means (Variogram.py:1769-1793)
            from scipy import stats
            from scipy.optimize import minimize
            weights = np.ones(_x.size) if sig is None else 1.0 / np.asarray(sig, dtype=float)

            def negative_loglik(params):
                pred = np.asarray([wrapped(h, *params) for h in _x], dtype=float)
                return -np.sum(stats.norm.logpdf(pred, loc=_y, scale=1.0) * weights)

            p0 = np.asarray(p0, dtype=float)
            res = minimize(negative_loglik, p0 * 0.5, method="SLSQP", bounds=[(0, b) for b in p0])
            if not res.success:
                raise RuntimeError("Maximum Likelihood could not estimate parameters.")
            self.cof = res.x

    def transform(self, x):
        if self.cof is None:
            self.fit(force=True)
        return self.fitted_model(x)

    @property
    def fitted_model(self):
        if self.cof is None:
            self.fit(force=True)
        return self.fitted_model_function(self._model, self.cof)

    @classmethod
    def fitted_model_function(cls, model, cof=None, **kw):
        """Variogram.py:1870-1900: cof = [range, sill, (shape), (nugget if != 0)]."""
        if cof is None:
            cof = [kw["effective_range"], kw["sill"]]
            if "smoothness" in kw:
                cof.append(kw["smoothness"])
            if "shape" in kw:
                cof.append(kw["shape"])
            if kw["nugget"] != 0:
                cof.append(kw["nugget"])
        if not callable(model):
            model = models.MODELS[str(model).lower()]
        cof = list(cof)

        def fitted_model(x):
            return model(x, *cof)

        return fitted_model

    def describe(self, short=False, flat=False):
        """Variogram.py:2627-2758 (single model)."""
        if self.cof is None:
            self.fit(force=True)
        if self.cof is None:
            raise RuntimeError("the variogram is not fitted (fit_method=None)")
        maxlag = np.nanmax(self.bins)
        maxvar = np.nanmax(self.experimental)
        cof = self.cof
        rdict = dict(
            model=self._model_name,
            estimator=self._estimator_name or getattr(self._estimator, "__name__", "custom"),
            dist_func=self.dist_function,
            normalized_effective_range=cof[0] * maxlag,
            normalized_sill=cof[1] * maxvar,
            normalized_nugget=cof[-1] * maxvar if self.use_nugget else 0,
            effective_range=cof[0],
            sill=cof[1],
            nugget=cof[-1] if self.use_nugget else 0,
        )
        if self._model_name == "matern":
            rdict["smoothness"] = cof[2]
        elif self._model_name == "stable":
            rdict["shape"] = cof[2]
        if not short:
            params = dict(
                estimator=rdict["estimator"], model=self._model_name, dist_func=str(self.dist_function),
                bin_func=self._bin_func_name, normalize=self.normalized,
                fit_method=self._fit_method, fit_sigma=self._fit_sigma,
                use_nugget=self.use_nugget, maxlag=self.maxlag,
                n_lags=self._n_lags_passed_value, verbose=self.verbose)
            if flat:
                rdict.update(params)
                rdict.update(self._kwargs)
            else:
                rdict["params"] = params
                rdict["kwargs"] = self._kwargs
        return rdict

    @property
    def parameters(self):
        d = self.describe()
        if self._model_name == "matern":
            return [d["effective_range"], d["sill"], d["smoothness"], d["nugget"]]
        if self._model_name == "stable":
            return [d["effective_range"], d["sill"], d["shape"], d["nugget"]]
        return [d["effective_range"], d["sill"], d["nugget"]]

    def data(self, n=100, force=False):
        """Theoretical variogram on n lags in [0, max(bins)] (Variogram.py:2187-2241)."""
        if self.cof is None or force:
            self.fit(force=True)
        x = np.linspace(0, np.nanmax(self.bins), n)
        return x, self.fitted_model(x)

    # --------------------------------------------------------- quality measures
    @property
    def model_residuals(self):
        """experimental - model at the lag edges (Variogram.py:2243-2268)."""
        exp = self.experimental
        return np.fromiter(map(lambda a, b: a - b, self.fitted_model(self.bins), exp), float)

    residuals = model_residuals

    @property
    def rmse(self):
        """Variogram.py:2300-2321: the reference returns root_mean_square (nanmean over the lag classes
        that hold pairs)."""
        return self.root_mean_square

    @property
    def mse(self):
        """Variogram.py:2326-2352: np.mean - an empty lag class (NaN) makes it NaN, as upstream."""
        experimental, model = self.model_deviations()
        return np.mean(np.power(np.subtract(experimental, model), 2))

    @property
    def mae(self):
        """Variogram.py:2358-2384: np.mean, NaN with an empty lag class."""
        experimental, model = self.model_deviations()
        return np.mean(np.abs(np.subtract(experimental, model)))

    @property
    def nrmse(self):
        return self.rmse / np.nanmean(self.experimental)

    @property
    def mean_residual(self):
        """Variogram.py:2270-2290."""
        return np.nanmean(np.fromiter(map(np.abs, self.model_residuals), float))

    @property
    def root_mean_square(self):
        """Variogram.py:2328-2343."""
        return np.sqrt(np.nanmean(np.square(self.model_residuals)))

    @property
    def residual_sum_of_squares(self):
        """Variogram.py:2345-2359."""
        return np.nansum(np.square(self.model_residuals))

    rss = residual_sum_of_squares

    @property
    def nrmse_r(self):
        """Variogram.py:2407-2436: rmse / (max - mean) of the experimental variogram."""
        y = self.experimental
        return self.rmse / (np.nanmax(y) - np.nanmean(y))

    def model_deviations(self):
        """(experimental, model at the bin edges), Variogram.py:2438-2470."""
        exp = self.experimental
        if self.cof is None:
            self.fit(force=True)
        if self.cof is None:
            raise RuntimeError("The variogram cannot be calculated.")
        return exp, self.transform(self.bins)

    @property
    def r(self):
        """Pearson correlation of experimental and model values, Variogram.py:2472-2513."""
        experimental, model = self.model_deviations()
        mx, my = np.nanmean(experimental), np.nanmean(model)
        term1 = np.nansum(np.fromiter(map(lambda x, y: (x - mx) * (y - my), experimental, model), float))
        t2x = np.nansum(np.fromiter(map(lambda x: (x - mx) ** 2, experimental), float))
        t2y = np.nansum(np.fromiter(map(lambda y: (y - my) ** 2, model), float))
        return term1 / (np.sqrt(t2x * t2y))

    pearson = r

    @property
    def NS(self):
        """Nash-Sutcliffe efficiency of the model, Variogram.py:2515-2553."""
        experimental, model = self.model_deviations()
        mx = np.nanmean(experimental)
        term1 = np.nansum(np.fromiter(map(lambda x, y: (x - y) ** 2, experimental, model), float))
        term2 = np.nansum(np.fromiter(map(lambda x: (x - mx) ** 2, experimental), float))
        return 1 - (term1 / term2)

    def to_DataFrame(self, n=100, force=False):
        """Variogram.py:2154-2185."""
        from pandas import DataFrame
        warnings.warn("The return value of this function will change in a future release.",
                      FutureWarning)
        lags, data = self.data(n=n, force=force)
        return DataFrame({"lags": lags, self._model_name: data}).copy()

    def update_kwargs(self, **kwargs):
        """Variogram.py:1478-1496."""
        self._kwargs.update(kwargs)

    def cross_validate(self, method="jacknife", n=None, metric="rmse", seed=None):
        """Variogram.py:2577-2625: leave-one-out kriging at the sample locations (one
        batched kernel call, see util/cross_validation.py)."""
        from .util import cross_validation
        if method == "jacknife":
            return cross_validation.jacknife(self, n=n, metric=metric, seed=seed)
        raise AttributeError("A method '%s' is not implemented." % method)

    def clone(self):
        return copy.deepcopy(self)

    def __repr__(self):
        try:
            r, s, n = self.parameters[0], self.parameters[1], self.parameters[-1]
            return "< %s Semivariogram fitted to %d bins >" % (self._model_name, len(self.bins))
        except Exception:
            return "< Variogram (B200) >"
