"""Monte-Carlo uncertainty propagation (reference: skgstat/util/uncertainty.py:12-198).

Same arguments, same random stream (numpy Generator seeded with `seed`, one draw per source and
iteration in the reference's order) and same confidence-interval arithmetic.  What changes is the
work per iteration: the reference constructs a full Variogram per draw (distance matrix lookups,
binning, estimator); here the draws that only perturb ``values`` share the coordinates, bins and
mask of ONE template variogram, and their experimental variograms come out of the batched pair
sweep (Variogram.experimental_batch -> skg_pair_hist_batch, four draws per sweep).
"""
from __future__ import annotations

import copy

import numpy as np

from ..Variogram import Variogram


def _evaluate(vario, evalf, eval_at):
    out = []
    if "experimental" in evalf:
        out.append(vario.experimental)
    if "parameter" in evalf:
        out.append(vario.parameters)
    if "model" in evalf:
        x = np.linspace(0, np.max(vario.bins), eval_at)
        out.append(vario.fitted_model(x))
    return out


def propagate(variogram=None, source="values", sigma=5, evalf="experimental", verbose=False,
              use_bounds=False, **kwargs):
    """skgstat/util/uncertainty.py:12-198: [lower, median, upper] of the evaluated property
    under random perturbation of `source` (per evalf entry; a single array if only one)."""
    if use_bounds:
        kwargs["q"] = 0
    metric_space = variogram._X
    if isinstance(source, str):
        source = [source]
    if not isinstance(sigma, (list, tuple)):
        sigma = [sigma]
    if isinstance(evalf, str):
        evalf = [evalf]
    var_opts = variogram.describe().get("params", {})
    omit = [*source, "verbose"]
    args = {k: v for k, v in var_opts.items() if k not in omit}
    args["coordinates"] = metric_space

    num_iter = kwargs.get("num_iter", 500)
    rng = np.random.default_rng(kwargs.get("seed"))
    dist = getattr(rng, kwargs.get("distribution", "normal"))
    param_field = []
    for _ in range(num_iter):
        par = {**args}
        for s, err in zip(source, sigma):
            obs = getattr(variogram, s)
            size = len(obs) if hasattr(obs, "__len__") else 1
            par[s] = dist(obs, err, size=size)
        param_field.append(par)

    eval_at = kwargs.get("eval_at", 100)
    result = None
    if source == ["values"] and num_iter > 0 and np.ndim(param_field[0]["values"]) == 1:
        # fast path: only the observations change -> one template, batched sweeps
        template = Variogram(**param_field[0])
        if template._estimator_name in ("matheron", "cressie", "dowd"):
            draws = np.vstack([par["values"] for par in param_field])
            exps = template.experimental_batch(draws)
            counts = template.bin_count
            need_fit = any(e in evalf for e in ("parameter", "model"))
            result = []
            for it in range(num_iter):
                v = copy.copy(template)
                v._values = draws[it]
                flat = draws[it]
                v._single_input = bool(np.all(flat == flat[0]))
                v._exp = exps[it]
                v._bin_count = counts
                v._diff = None
                v._values_token = (id(v), "draw", it)  # never the template's device values
                v.cof = None
                v.cov = None
                if need_fit:
                    v.fit(force=False)  # force would drop the injected per-draw statistics
                result.append(_evaluate(v, evalf, eval_at))
    if result is None:  # generic path: the reference's loop, each draw a full (GPU) Variogram
        result = [_evaluate(Variogram(**par), evalf, eval_at) for par in param_field]

    conf_intervals = []
    n_out = len(result[0]) if result else 0
    for i in range(n_out):
        res = [result[j][i] for j in range(len(result))]
        ql = int(kwargs.get("q", 10) / 2)
        qu = 100 - int(kwargs.get("q", 10) / 2)
        conf_intervals.append(np.column_stack((
            np.percentile(res, ql, axis=0),
            np.median(res, axis=0),
            np.percentile(res, qu, axis=0))))
    if len(conf_intervals) == 1:
        return conf_intervals[0]
    return conf_intervals
