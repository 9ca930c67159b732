"""DirectionalVariogram: the reference's API (DirectionalVariogram.py:23-346) with the
directional mask fused into every pair sweep.

The reference materialises per-pair angles with two O(N^2) Python-callback
``pdist`` passes (DirectionalVariogram.py:404, 409) and a boolean mask vector; here
the same decision (``_triangle`` :659-714 / ``_compass`` :749-782) is evaluated
per pair inside the kernels (csrc/common.cuh: dir_mask), so masked maxima,
masked order statistics (uniform bins), masked histograms and masked Dowd medians
are all single sweeps.  ``bandwidth='qNN'`` is the NN-th percentile of ALL
distances (DirectionalVariogram.py:503-509), found by the exact select passes.
"""
from __future__ import annotations

import numpy as np

from .engine import dir_params
from .MetricSpace import MetricSpace
from .Variogram import Variogram


class DirectionalVariogram(Variogram):
    def __init__(self, coordinates=None, values=None, estimator="matheron", model="spherical",
                 dist_func="euclidean", bin_func="even", normalize=False, fit_method="trf",
                 fit_sigma=None, directional_model="triangle", azimuth=0, tolerance=45.0,
                 bandwidth="q33", use_nugget=False, maxlag=None, n_lags=10, verbose=False,
                 **kwargs):
        if not isinstance(coordinates, MetricSpace):
            c = np.asarray(coordinates)
            if c.ndim != 2 or c.shape[1] != 2:
                raise NotImplementedError("N-dimensional coordinates cannot be handled")
        elif coordinates.coords.shape[1] != 2:
            raise NotImplementedError("N-dimensional coordinates cannot be handled")
        # directional settings must exist before the base constructor runs the first sweep
        self._azimuth = None
        self._tolerance = None
        self._bandwidth = None
        self._bandwidth_arg = bandwidth
        self._directional_model = None
        self._dir_ready = False
        self._pending = dict(azimuth=azimuth, tolerance=tolerance, bandwidth=bandwidth,
                             directional_model=directional_model)
        super().__init__(coordinates=coordinates, values=values, estimator=estimator,
                         model=model, dist_func=dist_func, bin_func=bin_func,
                         normalize=normalize, fit_method=fit_method, fit_sigma=fit_sigma,
                         use_nugget=use_nugget, maxlag=maxlag, n_lags=n_lags, verbose=verbose,
                         **kwargs)

    # The base constructor ends with fit(); directional parameters are installed right
    # before the first preprocessing so that one pass serves the constructor.
    def preprocessing(self, force=False):
        if not self._dir_ready:
            p = self._pending
            self._dir_ready = True
            self.azimuth = p["azimuth"]
            self.tolerance = p["tolerance"]
            self.bandwidth = p["bandwidth"]
            self.set_directional_model(p["directional_model"])
        super().preprocessing(force=force)

    # ---------------------------------------------------------------- settings
    def _invalidate_direction(self):
        if self._bin_func is not None and not callable(self._bin_func):
            self._bins = None  # masked edges depend on the mask (reference recomputes lazily)
        self._reset_pair_state()

    @property
    def azimuth(self):
        return self._azimuth

    @azimuth.setter
    def azimuth(self, angle):
        if angle < -180 or angle > 180:
            raise ValueError("The azimuth is an angle in degree and has to meet "
                             "-180 <= angle <= 180")
        self._azimuth = angle
        self._invalidate_direction()

    @property
    def tolerance(self):
        return self._tolerance

    @tolerance.setter
    def tolerance(self, angle):
        if angle < 0 or angle > 360:
            raise ValueError("The tolerance is an angle in degree and has to meet "
                             "0 <= angle <= 360")
        self._tolerance = angle
        self._invalidate_direction()

    @property
    def bandwidth(self):
        return 0 if self._bandwidth is None else self._bandwidth

    @bandwidth.setter
    def bandwidth(self, width):
        """DirectionalVariogram.py:503-520."""
        if isinstance(width, str):
            q = int(width[1:])
            self._bandwidth = self._distance_percentile(q)
        elif width < 0:
            raise ValueError("The bandwidth cannot be negative.")
        elif width > self._distance_max():
            print("The bandwidth is larger than the maximum separating distance. "
                  "Thus it will have no effect.")
        else:
            self._bandwidth = width
        self._bandwidth_arg = width
        self._invalidate_direction()

    def set_directional_model(self, model_name):
        if isinstance(model_name, str):
            name = model_name.lower()
            if name == "circle":
                raise NotImplementedError
            if name not in ("triangle", "compass"):
                raise ValueError("%s is not a valid model." % model_name)
            self._directional_model = name
        elif callable(model_name):
            raise NotImplementedError("custom directional search areas are evaluated on "
                                      "materialised angles in the reference and are outside "
                                      "the fused path")
        else:
            raise ValueError("The directional model has to be identified by a model name, "
                             "or it has to be the search area itself")
        self._invalidate_direction()

    # ------------------------------------------------------------ kernel hooks
    def _dirp(self):
        if not self._dir_ready or self._directional_model is None:
            return None
        return dir_params(self._directional_model, self._azimuth, self._tolerance,
                          self.bandwidth)

    def _dir_key(self):
        if not self._dir_ready:
            return None
        return (self._directional_model, float(self._azimuth), float(self._tolerance),
                float(self.bandwidth))

    def _direction_mask(self):
        """Boolean mask over the condensed pairs (DirectionalVariogram.py:605-624);
        materialising accessor."""
        out = self._engine().materialize(("mask",), dirp=self._dirp())
        return out["mask"].astype(bool)

    def to_gstools(self, *args, **kwargs):
        raise NotImplementedError("Exporting DirectinalVariogram is currently not supported.")
