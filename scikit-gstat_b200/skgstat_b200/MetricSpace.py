"""MetricSpace: a point cloud plus the euclidean metric, resident on the GPU.

Mirrors the reference's constructor (MetricSpace.py:87-116).  The reference's
``dists`` builds the dense N x N matrix with scipy (MetricSpace.py:135-156) and
everything downstream reads it; here the pair kernels work from the coordinates
directly and ``dists`` only exists as a lazily materialised accessor (produced by
skg_pair_materialize, bit-identical to scipy's pdist/squareform).

``max_dist`` gives the reference's sparse semantics (only pairs with 0 < d <= max_dist exist)
through a pair filter on the engine.  Out of scope:
non-euclidean metrics, Probabalistic/RasterEquidistant samplers - they raise
NotImplementedError instead of silently taking another path.
"""
from __future__ import annotations

import numpy as np


class MetricSpace:
    def __init__(self, coords, dist_metric="euclidean", max_dist=None, dist_metric_kwargs={}):
        if dist_metric != "euclidean":
            raise NotImplementedError(
                "skgstat_b200 implements the euclidean metric only (got %r)" % (dist_metric,))
        c = np.asarray(coords)
        if c.ndim != 2:
            raise ValueError("coords has to be of shape (Npoints, Ndim)")
        self.coords = c.copy()
        self.dist_metric = dist_metric
        self.max_dist = max_dist
        self.dist_metric_kwargs = dict(dist_metric_kwargs)
        self._dists = None
        self._engine = None
        self._cache = {}

    # ---- device side --------------------------------------------------------
    def engine(self, group=None):
        """The PairEngine holding these coordinates in HBM (created on first use)."""
        if self._engine is None or self._engine.group is not group:
            from .engine import PairEngine
            self._engine = PairEngine(self.coords, group=group)
            # sparse space (MetricSpace.py:141-147): only pairs with 0 < d <= max_dist exist
            self._engine.set_pair_filter(self.max_dist)
            self._cache = {}
        return self._engine

    def cached(self, key, fn):
        """Memoise coordinate-only scalars (max distance, order statistics)."""
        if key not in self._cache:
            self._cache[key] = fn()
        return self._cache[key]

    @property
    def dists(self):
        """N x N distance matrix (materialising accessor): dense, or - with max_dist - the
        scipy.sparse CSR matrix of the pairs with 0 < d <= max_dist (MetricSpace.py:141-147)."""
        if self._dists is None and self.max_dist is not None:
            from scipy import sparse
            eng = self.engine()
            d = eng.materialize(("dist",), _raw=True)["dist"]
            keep = np.flatnonzero((d > 0) & (d <= self.max_dist))
            iu, ju = np.triu_indices(len(self), 1)
            i, j, v = iu[keep], ju[keep], d[keep]
            m = sparse.coo_matrix((np.concatenate((v, v)), (np.concatenate((i, j)), np.concatenate((j, i)))),
                                  shape=(len(self), len(self)))
            self._dists = m.tocsr()
        if self._dists is None:
            from scipy.spatial.distance import squareform  # layout helper only
            d = self.engine().materialize(("dist",))["dist"]
            self._dists = squareform(d, checks=False) if len(self) > 1 else np.zeros((len(self),) * 2)
        return self._dists

    def __len__(self):
        return len(self.coords)

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"] = None  # device buffers / ctypes handles are not picklable
        st["_cache"] = {}
        return st


class MetricSpacePair:
    """Pair of point clouds (targets, samples) as OrdinaryKriging.transform builds it
    (MetricSpace.py:206-256).  Only a container here: the M x N distance matrix the
    reference materialises is never formed."""

    def __init__(self, ms1, ms2):
        if ms1.dist_metric != ms2.dist_metric:
            raise ValueError("Both MetricSpaces need to have the same distance metric")
        if ms1.max_dist != ms2.max_dist:
            raise ValueError("Both MetricSpaces need to have the same max_dist")
        self.ms1 = ms1
        self.ms2 = ms2

    @property
    def dist_metric(self):
        return self.ms1.dist_metric

    @property
    def max_dist(self):
        return self.ms1.max_dist
