"""OrdinaryKriging: the reference's API (Kriging.py:145-474) over the batched
warp-per-target kriging kernel (csrc/krige_kernels.cu, skg_krige).

``transform(*x)`` keeps the reference's contract: one 1-D array per coordinate
dimension (or a MetricSpace), returns the estimates, fills ``self.sigma`` with the
kriging variances, counts ``no_points_error`` / ``singular_error`` and emits the
same KrigingWarning summaries (Kriging.py:461-469).  Defaults follow the reference
(mode='exact' and 'estimate', solver='inv', n_jobs=1); sparse spaces and other
metrics are out of scope and raise NotImplementedError.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _lib, models
from .MetricSpace import MetricSpace, MetricSpacePair
from .Variogram import Variogram


class KrigingWarning(Warning):
    pass


class LessPointsError(RuntimeError):
    pass


class SingularMatrixError(np.linalg.LinAlgError):
    pass


class IllMatrixError(RuntimeWarning):
    pass


class OrdinaryKriging:
    def __init__(self, variogram, min_points=5, max_points=15, mode="exact", precision=100,
                 solver="inv", n_jobs=1, perf=False, sparse=False, coordinates=None, values=None,
                 process_group=None):
        if isinstance(variogram, Variogram):
            if coordinates is None:
                coordinates = variogram.coordinates
            if values is None:
                values = variogram.values
            variogram = variogram.describe()
        if sparse:
            raise NotImplementedError("sparse kriging spaces are out of scope")
        self.sparse = sparse
        self._minp = min_points
        self._maxp = max_points
        self.min_points = min_points
        self.max_points = max_points
        self.n_jobs = n_jobs
        self.perf = perf
        self.range = variogram["effective_range"]
        self.nugget = variogram["nugget"]
        self.sill = variogram["sill"]
        self._model_name = str(variogram["model"]).lower()
        if self._model_name not in models.DEVICE_MODELS:
            raise NotImplementedError("model %r is not implemented in the kriging kernel "
                                      "(available: %s)" % (self._model_name,
                                                           ", ".join(models.DEVICE_MODELS)))
        self._shape = variogram.get("smoothness") if self._model_name == "matern" \
            else variogram.get("shape")
        self.dist_metric = variogram.get("dist_func", "euclidean")
        if isinstance(coordinates, MetricSpace):
            assert self.dist_metric == coordinates.dist_metric, \
                "Distance metric of variogram differs from distance metric of coordinates"
            coords = coordinates
            vals = np.asarray(values)
        else:
            if self.dist_metric != "euclidean":
                raise NotImplementedError("only the euclidean metric is implemented")
            c, vals = self._remove_duplicated_coordinates(np.asarray(coordinates),
                                                          np.asarray(values))
            coords = MetricSpace(c.copy(), self.dist_metric, None)
        self.values = vals.copy()
        self.coords = coords
        self.gamma_model = Variogram.fitted_model_function(**variogram)
        self.z = None
        self.sigma = None
        # calculation mode; self.range has to be initialized (Kriging.py:259-265)
        self._mode = mode
        self._precision = precision
        self._prec_dist = None
        self._prec_g = None
        self.mode = mode
        self.precision = precision
        self._solver = None
        self.solver = solver
        self.singular_error = 0
        self.no_points_error = 0
        self.ill_matrix = 0
        self._process_group = process_group
        self._krige_engine = None

    @classmethod
    def _remove_duplicated_coordinates(cls, coords, values):
        """Keep the first occurrence of duplicated sample locations (Kriging.py:285-303)."""
        _, first = np.unique(coords, axis=0, return_index=True)
        first.sort()
        return coords[first], values[first]

    # ---- argument validation: same messages as the reference (Kriging.py:305-381) ----
    @property
    def min_points(self):
        return self._minp

    @min_points.setter
    def min_points(self, value):
        if not isinstance(value, int):
            raise ValueError("min_points has to be an integer.")
        if value < 0:
            raise ValueError("min_points can't be negative.")
        if value > self._maxp:
            raise ValueError("min_points can't be larger than max_points.")
        self._minp = value

    @property
    def max_points(self):
        return self._maxp

    @max_points.setter
    def max_points(self, value):
        if not isinstance(value, int):
            raise ValueError("max_points has to be an integer.")
        if value < 0:
            raise ValueError("max_points can't be negative.")
        if value < self._minp:
            raise ValueError("max_points can't be smaller than min_points.")
        self._maxp = value

    @property
    def mode(self):
        return self._mode

    @mode.setter
    def mode(self, value):
        """Kriging.py:343-352."""
        if value == "exact":
            self._prec_g = None
            self._prec_dist = None
        elif value == "estimate":
            self._precalculate_matrix()
        else:
            raise ValueError("mode has to be one of 'exact', 'estimate'.")
        self._mode = value
        self._krige_lut = None

    @property
    def precision(self):
        return self._precision

    @precision.setter
    def precision(self, value):
        if not isinstance(value, int):
            raise TypeError("precision has to be of type int")
        if value < 1:
            raise ValueError("The precision has be be > 1")
        self._precision = value
        self._precalculate_matrix()
        self._krige_lut = None

    def _precalculate_matrix(self):
        """Kriging.py:656-661: semivariances at `precision` distances in [0, range]; in
        'estimate' mode the kernel reads the kriging MATRIX entries from this table."""
        self._prec_dist = np.linspace(0, self.range, self._precision)
        self._prec_g = np.asarray(self.gamma_model(self._prec_dist), dtype=np.float64)

    @property
    def solver(self):
        return self._solver

    @solver.setter
    def solver(self, value):
        if value not in ("inv", "numpy", "scipy"):
            raise AttributeError("solver has to be ['inv', 'numpy', 'scipy']")
        self._solver = value  # all three map onto the kernel's partial-pivot LU

    # --------------------------------------------------------------- the path
    def _engine(self):
        if self._krige_engine is None:
            from .engine import KrigeEngine
            self._krige_engine = KrigeEngine(self.coords.coords, self.values, self.range,
                                             group=self._process_group)
        return self._krige_engine

    def _model_params(self):
        return np.array([self.range, self.sill,
                         self._shape if self._shape is not None else 0.0,
                         self.nugget if self.nugget != 0 else 0.0,
                         self.range, self.sill], dtype=np.float64)  # table range / beyond-table value

    def _matrix_lut(self):
        """mode='estimate': the precalculated semivariances (host array) or None."""
        return self._prec_g if self._mode == "estimate" else None

    def transform(self, *x):
        """Kriging.py:405-474."""
        self.singular_error = 0
        self.no_points_error = 0
        self.ill_matrix = 0
        # the per-dimension arrays go to the device as they are (SoA); the MetricSpace objects the
        # reference builds here (Kriging.py:430-440) are formed only if somebody reads them
        self._transform_space = None
        if len(x) == 1 and isinstance(x[0], MetricSpace):
            self._transform_space = x[0]
            t = np.ascontiguousarray(np.asarray(x[0].coords, dtype=np.float64).T)
        else:
            cols = [np.asarray(xi, dtype=np.float64).ravel() for xi in x]
            if len({c.size for c in cols}) > 1:
                raise ValueError("all input arrays must have the same length")
            t = np.vstack(cols) if cols else np.zeros((0, 0))
        self._transform_x = t
        if t.shape[0] != self.coords.coords.shape[1]:
            raise ValueError("transform coordinates have %d dimensions, the samples %d"
                             % (t.shape[0], self.coords.coords.shape[1]))
        if self._maxp > _lib.KRIGE_MAX_POINTS:
            raise NotImplementedError("max_points > %d is not supported by the warp-per-target "
                                      "kernel yet" % _lib.KRIGE_MAX_POINTS)
        z, sigma, status = self._engine().transform(t, self._model_name, self._model_params(),
                                                    self._minp, self._maxp,
                                                    matrix_lut=self._matrix_lut(), soa=True)
        self.no_points_error = int(np.count_nonzero(status == _lib.KRIGE_LESS_POINTS))
        self.singular_error = int(np.count_nonzero(status == _lib.KRIGE_SINGULAR))
        self.sigma = sigma
        if self.singular_error > 0:
            warnings.warn("%d kriging matrices were singular." % self.singular_error,
                          KrigingWarning)
        if self.no_points_error > 0:
            warnings.warn("For %d locations, not enough neighbors were found within the range."
                          % self.no_points_error, KrigingWarning)
        self.z = z
        return np.array(z)

    @property
    def transform_coords(self):
        """MetricSpace of the last transform's targets (Kriging.py:430-436), built on demand."""
        if getattr(self, "_transform_space", None) is None and getattr(self, "_transform_x", None) is not None:
            self._transform_space = MetricSpace(np.ascontiguousarray(self._transform_x.T), self.dist_metric, None)
        return getattr(self, "_transform_space", None)

    @property
    def transform_coords_pair(self):
        ms = self.transform_coords
        return None if ms is None else MetricSpacePair(ms, self.coords)

    def __call__(self, *x):
        return self.transform(*x)

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_krige_engine"] = None
        st["_process_group"] = None
        return st
