"""PairEngine: device-resident point set + the pair-sweep entry points.

Host orchestration only.  Every O(N^2) step runs in libskg_b200.so (hand-written
sm_100a kernels) through the ctypes C ABI; torch provides device buffers,
streams and (for N GPUs) the NCCL collectives.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import distributed as D
from ._exact import INF_BITS, from_bits, sq_threshold_bits, to_bits
from .select import exact_select


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("skgstat_b200 needs a CUDA device (B200 / sm_100a); "
                           "there is no CPU fallback for the hot path")


class Traffic:
    """Host<->device byte counters of the current process (bench.py reads them)."""
    h2d = 0
    d2h = 0

    @classmethod
    def reset(cls):
        cls.h2d = 0
        cls.d2h = 0


_PINNED = {}


def _pinned_like(t: torch.Tensor) -> torch.Tensor:
    """Reusable pinned staging buffer (cudaHostAlloc per upload would dominate small calls)."""
    key = (t.dtype, max(1, t.numel()))
    buf = _PINNED.get(key)
    if buf is None:
        cap = 1 << max(12, int(t.numel() - 1).bit_length()) if t.numel() > 1 else 4096
        key2 = (t.dtype, cap)
        buf = _PINNED.get(key2)
        if buf is None:
            try:
                buf = torch.empty(cap, dtype=t.dtype).pin_memory()
            except RuntimeError:
                buf = torch.empty(cap, dtype=t.dtype)
            _PINNED[key2] = buf
        _PINNED[key] = buf
    return buf


def to_device(a: np.ndarray, device) -> torch.Tensor:
    """Host array -> device through (reused) pinned memory; returns a new device tensor."""
    t = torch.from_numpy(np.ascontiguousarray(a))
    Traffic.h2d += t.numel() * t.element_size()
    if t.numel() == 0:
        return t.to(device)
    buf = _pinned_like(t)
    stage = buf[: t.numel()].view(t.shape) if t.dim() > 1 else buf[: t.numel()]
    stage.copy_(t)
    return stage.to(device)  # synchronous w.r.t. the host: the staging buffer is reusable


def to_host(t: torch.Tensor) -> np.ndarray:
    Traffic.d2h += t.numel() * t.element_size()
    return t.cpu().numpy()


import contextlib
import os as _os
import time as _time

_DEBUG_TIMES = bool(_os.environ.get("SKG_SELECT_DEBUG"))


@contextlib.contextmanager
def _phase(label):
    """Developer timing of one host-visible phase (SKG_SELECT_DEBUG=1); a no-op otherwise."""
    if not _DEBUG_TIMES:
        yield
        return
    torch.cuda.synchronize()
    t0 = _time.perf_counter()
    yield
    torch.cuda.synchronize()
    print("   [phase] %-28s %.3f ms" % (label, (_time.perf_counter() - t0) * 1e3), flush=True)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _hptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dir_params(model: str, azimuth_deg: float, tolerance_deg: float, bandwidth: float,
               exclude_ambiguous: bool = True):
    """[model id, np.radians(azimuth), np.radians(tolerance/2), bandwidth/2, policy] - the
    quantities DirectionalVariogram._triangle/_compass compare against (:704-714).
    policy 1: guard-band pairs are resolved on the host (PairEngine.ambiguous_pairs)."""
    mid = {"triangle": _lib.DIR_TRIANGLE, "compass": _lib.DIR_COMPASS}[model]
    return np.array([mid, np.radians(azimuth_deg), np.radians(tolerance_deg / 2),
                     bandwidth / 2, 1.0 if exclude_ambiguous else 0.0], dtype=np.float64)


def literal_mask(dx, dy, dirp):
    """The reference's mask decision for explicit pair vectors, with the host's numpy
    (DirectionalVariogram.py:398-413 angles, :704-714 _triangle, :776-782 _compass).
    dx, dy = P_i - P_j (i < j).  Only used for the guard-band pairs."""
    with np.errstate(invalid="ignore", divide="ignore"):
        ed = np.sqrt(dx * dx + dy * dy)
        pos = np.arccos(dx / ed)
        ang = np.where(dy >= 0, pos, -pos)
        t = np.abs(ang + dirp[1])
        a = np.where(t > np.pi, t - np.pi, t)
        a = np.where(a > np.pi / 2, np.pi - a, a)
        ok = a <= dirp[2]
        if int(dirp[0]) == _lib.DIR_TRIANGLE:
            ok &= dirp[3] >= np.abs(ed * np.sin(t))
    return ok


class PairEngine:
    """Pair sweeps over one point set resident in HBM (SoA fp64)."""

    def __init__(self, coords, values=None, device=None, group=None, value_mode=_lib.VALUE_SCALAR):
        _require_cuda()
        self.lib = _lib.load()
        c = np.ascontiguousarray(np.asarray(coords, dtype=np.float64))
        if c.ndim != 2 or c.shape[1] not in (2, 3):
            raise ValueError("coords must be (N, 2) or (N, 3); got %r" % (c.shape,))
        self.n, self.dim = int(c.shape[0]), int(c.shape[1])
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.group = group
        self.rank, self.world = D.world(group)
        # sweeps run over a Z-order sorted copy; materialising accessors need the
        # caller's order (condensed index) and use an unsorted copy made on demand
        self._orig_coords = c
        self._sq_ub = None
        self._perm_host = None
        self._host_coords_sorted = None
        self._values_unsorted = None
        with torch.cuda.device(self.device):
            ct = np.ascontiguousarray(c.T)          # SoA, transposed once (upload and extents)
            self._coords_unsorted = to_device(ct, self.device).view(self.dim, self.n)
            if self.n > 64:
                with np.errstate(invalid="ignore"):
                    lo_raw = np.array([np.nanmin(r) for r in ct])
                    ext_raw = np.array([np.nanmax(r) for r in ct]) - lo_raw
                    lo = np.nan_to_num(lo_raw)
                    ext = np.nan_to_num(ext_raw)
                ub = float(np.sum(ext_raw * ext_raw)) * (1.0 + 1e-12)   # same value sq_upper_bound() forms
                self._sq_ub = ub if ub > 0.0 else 1.0
                keys = torch.empty(self.n, dtype=torch.int64, device=self.device)
                rc = self.lib.skg_morton_keys(_ptr(self._coords_unsorted), self.n, self.dim,
                                              _hptr(lo), _hptr(ext), _ptr(keys), self._stream())
                _lib.check(rc, "skg_morton_keys")
                self._perm_dev = torch.argsort(keys)  # keys < 2^63: signed order == unsigned order
                self.coords = self._coords_unsorted[:, self._perm_dev].contiguous().view(-1)
            else:
                self._perm_dev = None
                self.coords = self._coords_unsorted.reshape(-1)
            self._coords_unsorted = self._coords_unsorted.reshape(-1)
        self.values = None
        self._host_values_sorted = None
        self._orig_values = None
        self._dz_bound = 0.0
        self.nv, self.vmode = 1, _lib.VALUE_SCALAR
        if values is not None:
            self.set_values(values, value_mode)
        self.n_tiles = _lib.num_tiles(self.n)
        # interleaved tile shards: rank r sweeps tiles r, r+world, ... (balanced under culling)
        self.t0, self.t1, self.tstride = self.rank, self.n_tiles, self.world
        if self.t0 >= self.t1:
            self.t0 = self.t1 = 0
        self._scratch = None
        self._all_finite = None
        self.pair_filter = None  # (exclusive upper bound of s as bits, drop s == 0): sparse max_dist spaces
        self.max_dist = None
        self._buffers = {}
        self._sub_engine = None
        self._amb_cache = {}
        self.n_ambiguous = 0
        self.launches = 0  # kernels of this library enqueued (bench bookkeeping)

    # ------------------------------------------------------------------ utils
    @property
    def n_pairs(self) -> int:
        return self.n * (self.n - 1) // 2

    @property
    def perm(self) -> np.ndarray:
        """sorted position -> caller's index (host copy, fetched on first use)."""
        if self._perm_host is None:
            self._perm_host = np.arange(self.n) if self._perm_dev is None \
                else to_host(self._perm_dev).astype(np.int64)
        return self._perm_host

    @property
    def _host_coords(self) -> np.ndarray:
        if self._host_coords_sorted is None:
            self._host_coords_sorted = np.ascontiguousarray(self._orig_coords[self.perm])
        return self._host_coords_sorted

    @property
    def _host_values(self):
        if self.values is None:
            return None
        if self._host_values_sorted is None:
            self._host_values_sorted = np.ascontiguousarray(self._orig_values[self.perm])
        return self._host_values_sorted

    def set_values(self, values, value_mode=_lib.VALUE_SCALAR):
        """values: (N,) observations, or (N, k) columns with value_mode VALUE_CROSS (k = 2,
        |dz_0|*|dz_1|, Variogram.py:1979-1984) / VALUE_AGGREGATE (sqrt(sum dz_k^2), :1974-1977)."""
        v = np.asarray(values, dtype=np.float64)
        if v.ndim == 2 and value_mode == _lib.VALUE_SCALAR and v.shape[1] == 1:
            v = v[:, 0]
        if v.ndim == 1:
            if value_mode != _lib.VALUE_SCALAR:
                v = v[:, None]
        elif v.ndim != 2 or value_mode == _lib.VALUE_SCALAR:
            raise ValueError("values must be (N,) or (N, k) with a cross/aggregate value_mode")
        if v.shape[0] != self.n:
            raise ValueError("values length %d != number of points %d" % (v.shape[0], self.n))
        if v.ndim == 2:
            k = v.shape[1]
            if (value_mode == _lib.VALUE_CROSS and k != 2) or not (1 <= k <= _lib.MAX_VALUE_COLUMNS):
                raise ValueError("cross values need 2 columns, aggregate values 1..%d; got %d"
                                 % (_lib.MAX_VALUE_COLUMNS, k))
        v = np.ascontiguousarray(v)
        self._orig_values = v
        self._values_version = getattr(self, "_values_version", 0) + 1
        self.nv = 1 if v.ndim == 1 else int(v.shape[1])
        self.vmode = value_mode
        cols = np.ascontiguousarray(v.T).reshape(self.nv, self.n)        # device layout [k][n]
        dev = to_device(cols, self.device).view(self.nv, self.n)
        self._values_unsorted = dev.reshape(-1)
        self.values = self._values_unsorted if self._perm_dev is None \
            else dev[:, self._perm_dev].contiguous().view(-1)
        self._host_values_sorted = None
        # upper bound of the per-pair difference, formed with the kernels' operand order so that
        # monotone rounding keeps it >= every pair's value
        with np.errstate(invalid="ignore"):
            rng = [float(abs(np.nanmax(c) - np.nanmin(c))) if c.size else 0.0 for c in cols]
        rng = [r if r == r else 0.0 for r in rng]
        if value_mode == _lib.VALUE_CROSS:
            self._dz_bound = rng[0] * rng[1]
        elif value_mode == _lib.VALUE_AGGREGATE:
            acc = 0.0
            for r in rng:
                acc = acc + r * r
            self._dz_bound = float(np.sqrt(acc))
        else:
            self._dz_bound = rng[0]

    def _host_diff(self, i, j):
        """Per-pair difference of host-resolved pairs (sorted indices), kernels' arithmetic."""
        v = self._host_values
        if self.vmode == _lib.VALUE_CROSS:
            return np.abs(v[i, 0] - v[j, 0]) * np.abs(v[i, 1] - v[j, 1])
        if self.vmode == _lib.VALUE_AGGREGATE:
            acc = np.zeros(len(i))
            for k in range(self.nv):
                d = v[i, k] - v[j, k]
                acc = acc + d * d
            return np.sqrt(acc)
        return np.abs(v[i] - v[j])

    def _buffer(self, name, numel, dtype):
        """Persistent device scratch (avoids allocator round trips on every call)."""
        buf = self._buffers.get(name)
        if buf is None or buf.numel() < numel or buf.dtype != dtype:
            buf = torch.empty(max(int(numel), 1), dtype=dtype, device=self.device)
            self._buffers[name] = buf
        return buf[:numel]

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _hist_scratch(self, n_lags):
        need = int(self.lib.skg_pair_scratch_bytes(self.n, int(n_lags)))
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._scratch

    def last_sweep_info(self):
        """(surviving work items, candidate work items, pair slots per item) of the last
        sweep that used this engine's scratch (see include/skg_b200.h, tile culling)."""
        if self._scratch is None:
            return None
        c = self._scratch[:32].view(torch.int64).cpu().numpy()
        js = max(int(c[2]), 1)
        return int(c[0]), int(c[3]), _lib.TILE * (_lib.TILE // js)

    # ---------------------------------------------- sparse (max_dist) spaces
    def set_pair_filter(self, max_dist):
        """MetricSpace(max_dist=...) semantics (MetricSpace.py:141-147, Variogram.py:1210-1227): only
        pairs with 0 < d <= max_dist exist (the reference's sparse matrix drops coincident points).
        Every sweep then works on lag thresholds clipped to that range, see _filter_thr."""
        from ._exact import sq_upper_inclusive_bits
        self.max_dist = None if max_dist is None else float(max_dist)
        self.pair_filter = None if max_dist is None else (int(sq_upper_inclusive_bits(float(max_dist))), True)

    def _filter_thr(self, thr):
        """thr (bit patterns of the lag thresholds) -> (thr2, src): thr2 adds a leading class for
        s == 0 (discarded) and clips at the filter's upper bound; src[k] = class of thr2 that
        holds user class k, or -1 when the class lies entirely beyond the bound."""
        if self.pair_filter is None:
            return thr, None
        hi, drop_zero = self.pair_filter
        out = [1] if drop_zero else []
        src = np.full(thr.size, -1, dtype=np.int64)
        lower = 0
        for k in range(thr.size):
            if lower < hi:
                src[k] = len(out)
                out.append(min(int(thr[k]), hi))
            lower = int(thr[k])
        thr2 = np.array(out, dtype=np.uint64)
        if thr2.size > _lib.MAX_LAGS:
            raise ValueError("too many lag classes for a sparse (max_dist) space")
        return thr2, src

    @staticmethod
    def _unmap(arr2, src, fill=0):
        """Per-class results of the filtered thresholds back to the user's classes."""
        if src is None:
            return arr2
        arr2 = np.asarray(arr2)
        out = np.full((src.size,) + arr2.shape[1:], fill, dtype=arr2.dtype if fill == 0 else np.float64)
        ok = src >= 0
        out[ok] = arr2[src[ok]]
        return out

    def sq_upper_bound(self) -> float:
        """A value >= every squared pair distance (bounding-box diagonal, padded)."""
        if self._sq_ub is None:
            ct = np.ascontiguousarray(self._orig_coords.T)
            ext = np.array([np.nanmax(r) - np.nanmin(r) for r in ct]) if self.n else np.zeros(self.dim)
            ub = float(np.sum(ext * ext)) * (1.0 + 1e-12)
            self._sq_ub = ub if ub > 0.0 else 1.0
        return self._sq_ub

    # ------------------------------------------------- guard-band resolution
    AMB_CAPACITY = 1 << 24

    def ambiguous_pairs(self, dirp, all_tiles=False):
        """Pairs (of this rank's tiles) inside the guard band of the directional decision
        that PASS the reference's literal numpy formula.  The kernels exclude every
        guard-band pair (policy 1); their contributions are added by the callers.
        Returns (i, j) index arrays, or None for isotropic sweeps / policy 0."""
        if dirp is None or dirp[4] == 0.0:
            return None
        t0, t1, ts = (0, self.n_tiles, 1) if all_tiles else (self.t0, self.t1, self.tstride)
        key = tuple(float(x) for x in dirp) + (t0, t1, ts)
        if key in self._amb_cache:
            return self._amb_cache[key]
        cap = self.AMB_CAPACITY
        pi = torch.empty(cap, dtype=torch.int32, device=self.device)
        pj = torch.empty(cap, dtype=torch.int32, device=self.device)
        cnt = torch.zeros(1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            scratch = self._hist_scratch(1)
            rc = self.lib.skg_pair_dir_ambiguous(
                _ptr(self.coords), self.n, self.dim, _hptr(dirp), _ptr(pi), _ptr(pj), _ptr(cnt),
                cap, _ptr(scratch), scratch.numel(), t0, t1, ts, self._stream())
        _lib.check(rc, "skg_pair_dir_ambiguous")
        self.launches += 4 if t1 > t0 else 0
        got = int(cnt.item())
        # a rank-local raise would leave the other ranks waiting in the next collective: agree on
        # the worst count first, then every rank raises
        worst = got if (all_tiles or self.world == 1) else D.allreduce_max_int(got, self.device, self.group)
        if worst > cap:
            got = worst
            raise RuntimeError(
                "%d point pairs lie exactly on the directional decision boundary (more than "
                "%d): lattice data with a round azimuth/tolerance; not supported" % (got, cap))
        i = to_host(pi[:got]).astype(np.int64)
        j = to_host(pj[:got]).astype(np.int64)
        order = np.lexsort((j, i))  # deterministic regardless of atomics order
        i, j = i[order], j[order]
        # the reference orients every pair as P_a - P_b with a < b in the caller's order
        oi, oj = self.perm[i], self.perm[j]
        swap = oi > oj
        a = np.where(swap, j, i)
        b = np.where(swap, i, j)
        c = self._host_coords
        ok = literal_mask(c[a, 0] - c[b, 0], c[a, 1] - c[b, 1], dirp)
        self._amb_cache[key] = (a[ok], b[ok])
        self.n_ambiguous = got
        return self._amb_cache[key]

    def _amb_keys(self, amb):
        """(bits of s, |dz|) of host-resolved pairs, with the kernels' arithmetic."""
        i, j = amb
        c = self._host_coords
        dx = c[i, 0] - c[j, 0]
        dy = c[i, 1] - c[j, 1]
        s = dx * dx + dy * dy
        if self.values is not None:
            dz = self._host_diff(i, j)
        else:
            dz = np.zeros(i.size)
        return np.ascontiguousarray(s).view(np.uint64), dz

    # ------------------------------------------------------------- K2 (fused)
    def hist(self, edges: Sequence[float], stats: int = _lib.STAT_SUM_SQ, dirp=None,
             thr_bits: Optional[np.ndarray] = None):
        """Fused distance -> lag class -> estimator statistics for the given UPPER
        edges.  Returns (count int64[n], sum_sq float64[n], sum_sqrt float64[n])."""
        if self.values is None:
            raise ValueError("values not set")
        thr = sq_threshold_bits(edges) if thr_bits is None else np.asarray(thr_bits, np.uint64)
        thr, fsrc = self._filter_thr(thr)
        nl = int(thr.size)
        out = self._buffer("hist_out", 3 * nl, torch.float64).zero_()
        cnt = self._buffer("hist_cnt", nl, torch.int64).zero_()
        scratch = self._hist_scratch(nl)
        with torch.cuda.device(self.device):
            rc = self.lib.skg_pair_hist(
                _ptr(self.coords), _ptr(self.values), self.nv, self.vmode, self.n, self.dim, _hptr(dirp), _hptr(thr),
                nl, int(stats), None, _ptr(cnt), _ptr(out[nl:2 * nl]), _ptr(out[2 * nl:]),
                _ptr(scratch), scratch.numel(), self.t0, self.t1, self.tstride, self._stream())
        _lib.check(rc, "skg_pair_hist")
        self.launches += 5 if self.t1 > self.t0 else 0  # boxes, flags, compact, sweep, reduce
        out[:nl] = cnt.to(torch.float64)  # exact: counts < 2^53
        amb = self.ambiguous_pairs(dirp)
        if amb is not None and amb[0].size:
            s_bits, dz = self._amb_keys(amb)
            g = np.searchsorted(thr, s_bits, side="right")
            keep = g < nl
            extra = np.zeros(3 * nl)
            np.add.at(extra, g[keep], 1.0)
            if stats & _lib.STAT_SUM_SQ:
                np.add.at(extra, nl + g[keep], dz[keep] * dz[keep])
            if stats & _lib.STAT_SUM_SQRT:
                np.add.at(extra, 2 * nl + g[keep], np.sqrt(dz[keep]))
            out += to_device(extra, self.device)
        D.allreduce_sum_(out, self.group)  # the single collective of the fused pass
        host = to_host(out)
        return (self._unmap(host[:nl].astype(np.int64), fsrc), self._unmap(host[nl:2 * nl].copy(), fsrc),
                self._unmap(host[2 * nl:].copy(), fsrc))

    def dist_moments(self, limit: float, dirp=None, third: bool = True):
        """(n, sum d, sum d^2, sum d^3) over the (mask-passing) pairs with d <= limit: the
        inputs of np.std / the skewness inside np.histogram_bin_edges (binning.py:114-122).
        Fused-pass sweeps whose "value difference" is a power of the distance (SKG_VALUE_DIST_*)."""
        from ._exact import sq_upper_inclusive_bits
        thr = np.array([sq_upper_inclusive_bits(limit)], dtype=np.uint64)  # class 0: d <= limit
        thr, fsrc = self._filter_thr(thr)
        cls = 0 if fsrc is None else int(fsrc[0])   # the class that holds 0 < d <= limit
        ncl = int(thr.size)
        amb = self.ambiguous_pairs(dirp)

        def sweep(mode, stats):
            out = self._buffer("hist_out", 3 * ncl, torch.float64).zero_()
            cnt = self._buffer("hist_cnt", ncl, torch.int64).zero_()
            scratch = self._hist_scratch(ncl)
            with torch.cuda.device(self.device):
                rc = self.lib.skg_pair_hist(
                    _ptr(self.coords), None, 0, mode, self.n, self.dim, _hptr(dirp), _hptr(thr), ncl,
                    int(stats), None, _ptr(cnt), _ptr(out[ncl:2 * ncl]), _ptr(out[2 * ncl:]), _ptr(scratch),
                    scratch.numel(), self.t0, self.t1, self.tstride, self._stream())
            _lib.check(rc, "skg_pair_hist")
            self.launches += 5 if self.t1 > self.t0 else 0
            out[:ncl] = cnt.to(torch.float64)
            if amb is not None and amb[0].size:
                s_bits, _ = self._amb_keys(amb)
                sv = s_bits[(s_bits < thr[cls]) & (s_bits >= (thr[cls - 1] if cls else 0))].view(np.float64)
                d = np.sqrt(sv)
                val = {_lib.VALUE_DIST_S: sv, _lib.VALUE_DIST_D: d, _lib.VALUE_DIST_D15: d * np.sqrt(d)}[mode]
                extra = np.zeros(3 * ncl)
                extra[cls], extra[ncl + cls], extra[2 * ncl + cls] = sv.size, float(np.sum(val * val)), float(np.sum(np.sqrt(val)))
                out += to_device(extra, self.device)
            D.allreduce_sum_(out, self.group)
            h = to_host(out)
            return np.array([h[cls], h[ncl + cls], h[2 * ncl + cls]])

        a = sweep(_lib.VALUE_DIST_S, _lib.STAT_SUM_SQ | _lib.STAT_SUM_SQRT)   # sum d^4, sum d
        b = sweep(_lib.VALUE_DIST_D, _lib.STAT_SUM_SQ)                        # sum d^2
        s3 = float(sweep(_lib.VALUE_DIST_D15, _lib.STAT_SUM_SQ)[1]) if third else None
        return int(a[0]), float(a[2]), float(b[1]), s3

    # -------------------------------------------------------- K2b (batched)
    BATCH_CHUNK = 64  # value vectors uploaded per call

    def hist_batch(self, edges, values, stat: int = _lib.STAT_SUM_SQ, dirp=None,
                   thr_bits: Optional[np.ndarray] = None):
        """Estimator sums of MANY value vectors over the same coordinates and lag edges:
        values (K, N) -> (count int64[n_lags], sums float64[K, n_lags]) with
        stat = STAT_SUM_SQ (sum dz^2) or STAT_SUM_SQRT (sum sqrt|dz|).  Four vectors share one
        pair sweep (skg_pair_hist_batch); the engine's own values are left untouched."""
        V = np.asarray(values, dtype=np.float64)
        if V.ndim != 2 or V.shape[1] != self.n:
            raise ValueError("values must be (K, %d); got %r" % (self.n, V.shape))
        if stat not in (_lib.STAT_SUM_SQ, _lib.STAT_SUM_SQRT):
            raise ValueError("stat must be STAT_SUM_SQ or STAT_SUM_SQRT")
        thr = sq_threshold_bits(edges) if thr_bits is None else np.asarray(thr_bits, np.uint64)
        thr, fsrc = self._filter_thr(thr)
        nl = int(thr.size)
        K = int(V.shape[0])
        if nl > self.lib.skg_pair_batch_max_lags():
            raise ValueError("n_lags=%d exceeds the batched pass limit %d"
                             % (nl, self.lib.skg_pair_batch_max_lags()))
        amb = self.ambiguous_pairs(dirp)
        sums = np.empty((K, nl))
        counts = None
        need = int(self.lib.skg_pair_batch_scratch_bytes(self.n, nl))
        scratch = self._buffer("batch_scratch", need, torch.uint8)
        for k0 in range(0, K, self.BATCH_CHUNK):
            k1 = min(K, k0 + self.BATCH_CHUNK)
            kb = k1 - k0
            dev = to_device(V[k0:k1], self.device).view(kb, self.n)
            if self._perm_dev is not None:
                dev = dev[:, self._perm_dev].contiguous()
            out = torch.zeros((kb + 1) * nl, dtype=torch.float64, device=self.device)
            cnt = self._buffer("hist_cnt", nl, torch.int64).zero_()
            with torch.cuda.device(self.device):
                rc = self.lib.skg_pair_hist_batch(
                    _ptr(self.coords), _ptr(dev), kb, self.n, self.dim, _hptr(dirp), _hptr(thr), nl,
                    int(stat), _ptr(cnt), _ptr(out[nl:]), _ptr(scratch), scratch.numel(),
                    self.t0, self.t1, self.tstride, self._stream())
            _lib.check(rc, "skg_pair_hist_batch")
            self.launches += (3 + 2 * ((kb + 3) // 4)) if self.t1 > self.t0 else 0
            out[:nl] = cnt.to(torch.float64)
            if amb is not None and amb[0].size:  # guard-band pairs, host-resolved
                i, j = amb
                s_bits, _ = self._amb_keys(amb)
                g = np.searchsorted(thr, s_bits, side="right")
                keep = g < nl
                extra = np.zeros((kb + 1, nl))
                np.add.at(extra[0], g[keep], 1.0)
                perm = self.perm
                for q in range(kb):
                    vq = V[k0 + q][perm]
                    dz = np.abs(vq[i] - vq[j])[keep]
                    np.add.at(extra[1 + q], g[keep], dz * dz if stat == _lib.STAT_SUM_SQ else np.sqrt(dz))
                out += to_device(extra.ravel(), self.device)
            D.allreduce_sum_(out, self.group)
            host = to_host(out).reshape(kb + 1, nl)
            if counts is None:
                counts = host[0].astype(np.int64)
            sums[k0:k1] = host[1:]
        if fsrc is not None:
            sums = self._unmap(sums.T, fsrc).T
        return self._unmap(counts, fsrc), np.ascontiguousarray(sums)

    # ------------------------------------------------ K1/K3 (order statistics)
    def _select_runner(self, key_kind, thr, dirp, n_buckets, collect_max=None):
        nl = 0 if thr is None else int(thr.size)
        nseg = max(nl, 1)
        scratch = self._hist_scratch(1)
        amb = self.ambiguous_pairs(dirp)
        amb_seg = amb_key = None
        if amb is not None and amb[0].size:
            s_bits, dz = self._amb_keys(amb)
            amb_seg = np.searchsorted(thr, s_bits, side="right") if nl else np.zeros(s_bits.size, np.int64)
            amb_key = s_bits.view(np.float64) if key_kind == _lib.KEY_SQDIST else dz
            ok = (amb_seg < nseg) & np.isfinite(amb_key)
            amb_seg, amb_key = amb_seg[ok], np.ascontiguousarray(amb_key[ok])

        def amb_slots(lo, hi, sc):
            """(keys, slots) of the host-resolved pairs inside the current windows."""
            kb = amb_key.view(np.uint64)
            inside = (kb >= lo[amb_seg]) & (kb < hi[amb_seg])
            k, g = amb_key[inside], amb_seg[inside]
            t = (k - lo[g].view(np.float64)) * sc[g]
            b = np.minimum(n_buckets - 1, np.where(t >= 9.0e18, n_buckets - 1, t).astype(np.int64))
            return k, g * n_buckets + b

        def run_hist(lo, hi, sc):
            with _phase("hist sweep (%d x %d)" % (nseg, n_buckets)):
                return run_hist_(lo, hi, sc)

        def run_gather(lo, hi, sc, bitmap, cap, want=None):
            with _phase("gather sweep (cap %d)" % cap):
                return run_gather_(lo, hi, sc, bitmap, cap, want)

        def run_hist_(lo, hi, sc):
            hist = self._buffer("select_hist", nseg * n_buckets, torch.int64).zero_()
            mx = self._buffer("select_max", 1, torch.int64).zero_()
            lo = np.ascontiguousarray(lo, dtype=np.uint64)
            hi = np.ascontiguousarray(hi, dtype=np.uint64)
            sc = np.ascontiguousarray(sc, dtype=np.float64)
            with torch.cuda.device(self.device):
                rc = self.lib.skg_pair_select_hist(
                    _ptr(self.coords), _ptr(self.values), self.nv, self.vmode, self.n, self.dim, _hptr(dirp),
                    _hptr(thr), nl, key_kind, _hptr(lo), _hptr(hi), _hptr(sc), n_buckets,
                    _ptr(hist), _ptr(mx) if (nl == 0 and collect_max is not None) else None,
                    _ptr(scratch), scratch.numel(), self.t0, self.t1, self.tstride, self._stream())
            _lib.check(rc, "skg_pair_select_hist")
            self.launches += 4 if self.t1 > self.t0 else 0
            if amb_key is not None and amb_key.size:
                _, slots = amb_slots(lo, hi, sc)
                extra = np.bincount(slots, minlength=nseg * n_buckets).astype(np.int64)
                hist += to_device(extra, self.device)
                if nl == 0:
                    mx = torch.maximum(mx, torch.tensor([int(amb_key.view(np.int64).max())],
                                                        dtype=torch.int64, device=self.device))
            D.allreduce_sum_(hist, self.group)
            if nl == 0 and collect_max is not None:
                D.allreduce_max_(mx, self.group)
                collect_max.append(int(mx.item()))
            return to_host(hist)

        def run_gather_(lo, hi, sc, bitmap, cap, want=None):
            """want = [(slot, [ranks relative to the slot's keys])]: the order statistics are
            picked on the device (sort of the gathered keys, one small copy back) and
            {slot: (count, {rank: key})} is returned instead of the raw candidates."""
            cap = int(cap)
            keys = self._buffer("gather_keys", max(cap, 1), torch.float64)
            slots = self._buffer("gather_slots", max(cap, 1), torch.int32)
            cnt = self._buffer("gather_cnt", 1, torch.int64).zero_()
            # single rank, ranks picked on the device: no host round trip for the candidate count -
            # unwritten entries keep slot -1, which no wanted slot matches, and exact_select checks
            # the per-slot counts that come back with the picks
            nosync = self.world == 1 and want is not None and not (amb_key is not None and amb_key.size)
            if nosync:
                slots.fill_(-1)
            bm = to_device(bitmap.view(np.int32), self.device)
            lo = np.ascontiguousarray(lo, dtype=np.uint64)
            hi = np.ascontiguousarray(hi, dtype=np.uint64)
            sc = np.ascontiguousarray(sc, dtype=np.float64)
            with torch.cuda.device(self.device):
                rc = self.lib.skg_pair_select_gather(
                    _ptr(self.coords), _ptr(self.values), self.nv, self.vmode, self.n, self.dim, _hptr(dirp),
                    _hptr(thr), nl, key_kind, _hptr(lo), _hptr(hi), _hptr(sc), n_buckets,
                    _ptr(bm), _ptr(keys), _ptr(slots), _ptr(cnt), cap, _ptr(scratch),
                    scratch.numel(), self.t0, self.t1, self.tstride, self._stream())
            _lib.check(rc, "skg_pair_select_gather")
            self.launches += 4 if self.t1 > self.t0 else 0
            if not nosync:
                got = int(cnt.item())
                worst = got if self.world == 1 else D.allreduce_max_int(got, self.device, self.group)
                if worst > cap:   # decided collectively, so every rank raises (no rank is left in a collective)
                    raise RuntimeError("gather overflow: %d > %d" % (worst, cap))
                keys, slots = keys[:got], slots[:got]
            if amb_key is not None and amb_key.size:
                k, sl = amb_slots(lo, hi, sc)
                hit = ((bitmap[sl >> 5] >> (sl & 31).astype(np.uint32)) & 1).astype(bool)
                keys = torch.cat([keys, to_device(k[hit], self.device)])
                slots = torch.cat([slots, to_device(sl[hit].astype(np.uint32).view(np.int32),
                                                     self.device)])
            k = D.allgather_varlen(keys, self.group)
            s = D.allgather_varlen(slots, self.group)
            if want is None:
                return to_host(k), to_host(s).view(np.uint32).astype(np.int64)
            # all wanted order statistics with a constant number of device operations: sort by key,
            # stable sort by slot, segment starts from a searchsorted, one gather, one copy back
            if not want:
                return {}
            n_ranks = sum(len(rel) for _, rel in want)
            if len(want) <= 32 and n_ranks <= 64:
                # ranked on the device by our own kernels (skg_pick_order_stats): by counting for a
                # few thousand keys, after one narrowing histogram for more; heavy ties (-1) fall
                # through to the sort below
                w_slot = np.array([slot for slot, _ in want], dtype=np.uint32)
                w_off = np.zeros(len(want) + 1, dtype=np.int32)
                w_off[1:] = np.cumsum([len(rel) for _, rel in want])
                w_rank = np.array([r for _, rel in want for r in rel], dtype=np.int64)
                seg = (w_slot.astype(np.int64) // int(n_buckets)).astype(np.int64)
                w_lo = np.ascontiguousarray(lo[seg])
                w_hi = np.ascontiguousarray(hi[seg])
                picked = self._buffer("pick_out", 32 + 64, torch.float64)
                pscr = self._buffer("pick_scratch", self.lib.skg_pick_scratch_bytes(), torch.uint8)
                if nosync:
                    pk, ps, pc, pcap = keys, slots, cnt, cap
                else:   # the candidates of all ranks (and the guard-band pairs), count known on the host
                    pk, ps = k.contiguous(), s.contiguous()
                    pcap = int(pk.numel())
                    pc = self._buffer("pick_cnt", 1, torch.int64).fill_(pcap)
                with torch.cuda.device(self.device):
                    rc = self.lib.skg_pick_order_stats(
                        _ptr(pk), _ptr(ps), _ptr(pc), pcap, _hptr(w_slot), _hptr(w_off), _hptr(w_rank),
                        _hptr(w_lo), _hptr(w_hi), len(want), _ptr(picked), _ptr(pscr), pscr.numel(),
                        self._stream())
                _lib.check(rc, "skg_pick_order_stats")
                self.launches += 1 if pcap <= _lib.PICK_DIRECT_MAX else 4
                flat = to_host(picked[:len(want) + n_ranks])
                if not np.any(flat[:len(want)] < 0):
                    out, pos = {}, len(want)
                    for w, (slot, rel) in enumerate(want):
                        out[slot] = (int(flat[w]), {r: float(flat[pos + i]) for i, r in enumerate(rel)})
                        pos += len(rel)
                    return out
            slots_w = torch.tensor([int(np.array([slot], dtype=np.uint32).view(np.int32)[0]) for slot, _ in want],
                                   dtype=torch.int32, device=self.device)
            if k.numel():
                k_sorted, order = torch.sort(k)
                s_sorted, order2 = torch.sort(s[order], stable=True)
                k_sorted = k_sorted[order2]
                start = torch.searchsorted(s_sorted, slots_w, right=False)
                stop = torch.searchsorted(s_sorted, slots_w, right=True)
            else:
                k_sorted = torch.zeros(1, dtype=torch.float64, device=self.device)
                start = torch.zeros(len(want), dtype=torch.int64, device=self.device)
                stop = start
            cnt_slot = (stop - start)
            rel_flat, seg_of = [], []
            for w, (slot, rel) in enumerate(want):
                rel_flat += [int(r) for r in rel]
                seg_of += [w] * len(rel)
            if rel_flat:
                seg_t = torch.tensor(seg_of, dtype=torch.int64, device=self.device)
                rel_t = torch.tensor(rel_flat, dtype=torch.int64, device=self.device)
                idx = start[seg_t] + torch.minimum(rel_t, torch.clamp(cnt_slot[seg_t] - 1, min=0))
                idx = torch.clamp(idx, 0, max(int(k_sorted.numel()) - 1, 0))
                vals = k_sorted[idx]
            else:
                vals = torch.zeros(0, dtype=torch.float64, device=self.device)
            flat = to_host(torch.cat([cnt_slot.to(torch.float64), vals]))
            nw = len(want)
            out, pos = {}, nw
            for w, (slot, rel) in enumerate(want):
                out[slot] = (int(flat[w]), {r: float(flat[pos + i]) for i, r in enumerate(rel)})
                pos += len(rel)
            return out

        run_gather.device_pick = True
        return run_hist, run_gather

    def _buckets_for(self, n_keys, nseg):
        b = 1 << 12
        while b < (1 << 22) and b * 4096 < n_keys // max(nseg, 1):
            b <<= 1
        while nseg * b > (1 << 24) and b > (1 << 8):
            b >>= 1
        return b

    CELL_SHIFT = 11                 # 512 cells per octave of s: a cell spans 0.07 % of d
    DIRECT_GATHER_LIMIT = 1 << 22   # cells up to this many keys are gathered (and sorted on the device) without a fine histogram

    def sq_cell_counts(self, dirp, lo_bits: int, hi_bits: int):
        """Exact counts of the (mask-passing) pairs with lo_bits <= bits(s) < hi_bits on a
        log-spaced grid of s (skg_pair_cell_counts; shared-memory counters, no threshold
        compares).  Returns (upper cell boundaries as bit patterns, clipped to [lo, hi]; counts)."""
        nc, shift = _lib.CELL_COUNT, self.CELL_SHIFT
        top = ((hi_bits - 1) >> 32) >> shift
        key0 = int(top) - (nc - 1)
        cnt = self._buffer("cell_cnt", nc, torch.int64).zero_()
        need = int(self.lib.skg_pair_cell_scratch_bytes(self.n))
        scratch = self._buffer("cell_scratch", need, torch.uint8)
        with torch.cuda.device(self.device):
            rc = self.lib.skg_pair_cell_counts(
                _ptr(self.coords), self.n, self.dim, _hptr(dirp), int(lo_bits), int(hi_bits), shift,
                key0, _ptr(cnt), _ptr(scratch), scratch.numel(), self.t0, self.t1, self.tstride,
                self._stream())
        _lib.check(rc, "skg_pair_cell_counts")
        self.launches += 5 if self.t1 > self.t0 else 0
        if dirp is not None and dirp[4] != 0.0 and lo_bits == 0 and hi_bits > to_bits(self.sq_upper_bound()) \
                and self.pair_filter is None:
            # the sweep met every pair the mask can pass: no guard-band pair seen (on any rank) means
            # there is none, and the listing sweep (skg_pair_dir_ambiguous) is not needed at all
            seen = int(scratch[:64].view(torch.int64)[6].item())
            if D.allreduce_max_int(seen, self.device, self.group) == 0:
                empty = (np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64))
                for tiles in ((self.t0, self.t1, self.tstride), (0, self.n_tiles, 1)):
                    self._amb_cache.setdefault(tuple(float(x) for x in dirp) + tiles, empty)
        amb = self.ambiguous_pairs(dirp)
        if amb is not None and amb[0].size:
            s_bits, _ = self._amb_keys(amb)
            sb = s_bits[(s_bits >= np.uint64(lo_bits)) & (s_bits < np.uint64(hi_bits))]
            cell = np.clip(((sb >> np.uint64(32)) >> np.uint64(shift)).astype(np.int64) - key0, 0, nc - 1)
            cnt += to_device(np.bincount(cell, minlength=nc).astype(np.int64), self.device)
        D.allreduce_sum_(cnt, self.group)
        counts = to_host(cnt).astype(np.int64)
        keys = np.maximum(key0 + 1 + np.arange(nc, dtype=np.int64), 0).astype(np.uint64)
        thr = np.clip(keys << np.uint64(32 + shift), np.uint64(lo_bits), np.uint64(hi_bits))
        thr[nc - 1] = hi_bits
        return thr, counts

    def hist_counts(self, thr: np.ndarray, dirp=None) -> np.ndarray:
        """Exact pair counts per class of squared distance for bit-pattern thresholds
        `thr` (skg_pair_hist with no estimator statistic: private counters, no atomics).
        No values needed; all-reduced over the group."""
        thr, fsrc = self._filter_thr(np.asarray(thr, np.uint64))
        nl = int(thr.size)
        cnt = self._buffer("hist_cnt", nl, torch.int64).zero_()
        scratch = self._hist_scratch(nl)
        vals = self.values if self.values is not None else self._buffer("zeros", self.n, torch.float64)
        with torch.cuda.device(self.device):
            rc = self.lib.skg_pair_hist(
                _ptr(self.coords), _ptr(vals), 1, _lib.VALUE_SCALAR, self.n, self.dim, _hptr(dirp), _hptr(thr), nl, 0,
                None, _ptr(cnt), None, None, _ptr(scratch), scratch.numel(), self.t0, self.t1,
                self.tstride, self._stream())
        _lib.check(rc, "skg_pair_hist")
        self.launches += 5 if self.t1 > self.t0 else 0
        amb = self.ambiguous_pairs(dirp)
        if amb is not None and amb[0].size:
            s_bits, _ = self._amb_keys(amb)
            g = np.searchsorted(thr, s_bits, side="right")
            extra = np.bincount(g[g < nl], minlength=nl).astype(np.int64)
            cnt += to_device(extra, self.device)
        D.allreduce_sum_(cnt, self.group)
        return self._unmap(to_host(cnt).astype(np.int64), fsrc)

    SUBSAMPLE_SELECT_MIN_PAIRS = 10 ** 9   # below this the 4096-cell sweep is cheap enough
    SUBSAMPLE_POINTS = 8192
    REFINE_CELLS = 16                      # narrow cells counted around an interpolated rank
    REFINE_CELL_KEYS = 1 << 20             # keys aimed for per refinement cell
    _select_mode = "auto"                  # "cells": always the log-cell sweep (tests)

    def _subsample_engine(self):
        """The same fixed-seed sub-sample of the points on every rank, swept by every rank on its own
        (group LOCAL): the problem is tiny, its collectives and synchronisations would cost more than
        its sweeps, and identical inputs give identical windows on all ranks."""
        if self._sub_engine is None:
            m = min(self.n, self.SUBSAMPLE_POINTS)
            idx = np.sort(np.random.default_rng(777).choice(self.n, size=m, replace=False))
            self._sub_idx = idx
            self._sub_engine = PairEngine(self._orig_coords[idx], None, device=self.device, group=D.LOCAL)
        return self._sub_engine

    def _window_cell_counts(self, rank_fn, hi_bits: int, limit):
        """Cells of squared distance (upper bounds as bit patterns, exact pair counts) that are
        narrow around every wanted rank - the input sq_order_stats gathers from - found WITHOUT
        a sweep over all pairs: (1) the wanted quantiles of a fixed-seed sub-sample give a window
        per rank; (2) exact counts per cell come from the count-only fused pass, in which every
        work item / point group whose box lies inside one cell is counted in bulk (rows x columns),
        so only the pairs near a cell boundary are evaluated; (3) cells still too crowded to gather
        are subdivided and counted again.  A window that misses its rank just leaves a wide cell,
        which step (3) refines: the result is exact either way."""
        sub = self._subsample_engine()
        m = float(sub.n)
        state = {}

        def sub_ranks(t):
            # ranks of the window ends in the sub-sample: the wanted quantile -/+ 3 sigma of the
            # empirical pair-distance distribution of m points (a miss only costs a refinement)
            state["t"] = t
            state["lo"], state["hi"] = [], []
            if t < 1000:
                return []
            for r in sorted(set(int(r) for r in rank_fn(t))):
                f = (r + 0.5) / t
                h = 3.0 * np.sqrt(max(f * (1.0 - f), 1.0 / m) / m) + 4.0 / m
                state["lo"].append(max(0, int(np.floor((f - h) * t))))
                state["hi"].append(min(t - 1, int(np.ceil((f + h) * t))))
            return sorted(set(state["lo"] + state["hi"]))

        with _phase("sub-sample select"):
            vals, st = sub.sq_order_stats(sub_ranks, limit=limit)
        lo_r, hi_r = state["lo"], state["hi"]
        if not lo_r:
            return None
        self.launches += sub.launches
        sub.launches = 0
        windows = []
        for a, b in zip(lo_r, hi_r):
            wl = 0 if a == 0 else to_bits(vals[a])
            wh = hi_bits if b >= st - 1 else min(hi_bits, to_bits(vals[b]) + 1)
            if wh > wl:
                windows.append([wl, wh])
        windows.sort()
        merged = []
        for w in windows:
            if merged and w[0] <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], w[1])
            else:
                merged.append(w)
        if not merged or len(merged) > 2:
            # many quantiles at once (uniform-count bins): the windows of an 8192-point sample would
            # together hold a third of all pairs - the log-cell sweep is the better first pass there
            return None
        # cells: [0, w0.lo) | w0 | [w0.hi, w1.lo) | w1 | ... | [wK.hi, hi_bits).  Few, wide classes: the
        # count-only fused pass handles them with bulk counts and register windows; the fine
        # resolution inside the wanted windows comes from the histogram / gather passes, which only
        # look at the tiles those windows touch
        bounds, pos = [], 0
        for wl, wh in merged:
            if wl > pos:
                bounds.append(wl)
            bounds.append(wh)
            pos = wh
        if pos < hi_bits:
            bounds.append(hi_bits)
        thr = np.array(sorted(set(bounds)), dtype=np.uint64)
        with _phase("count sweep (%d classes)" % thr.size):
            counts = self.hist_counts(thr)
        # second level: inside a window the distribution of s is smooth, so linear interpolation of
        # the exact counts at its ends predicts a wanted rank to ~1 % of the window's keys; a band
        # of REFINE_CELLS narrow cells around the prediction is counted exactly (again in bulk
        # outside the band) and the cell that holds the rank is small enough to gather directly
        total = int(counts.sum())
        ranks = sorted(set(int(r) for r in rank_fn(total)))
        if not ranks:
            return thr, counts
        cum = np.cumsum(counts)
        lows = np.concatenate((np.zeros(1, dtype=np.uint64), thr[:-1]))   # (a Python 0 would promote to float64)
        bands = {}   # window class -> thresholds strictly inside it
        centres = {}  # window class -> relative positions already banded (adjacent ranks share one band)
        for r in ranks:
            c = int(np.searchsorted(cum, r, side="right"))
            n_in = int(counts[c])
            if n_in <= self.DIRECT_GATHER_LIMIT // 2:
                continue
            lo_v, hi_v = from_bits(int(lows[c])), from_bits(int(thr[c]) - 1)
            if not (np.isfinite(lo_v) and np.isfinite(hi_v) and hi_v > lo_v):
                continue
            rel = (r - (int(cum[c - 1]) if c else 0) + 0.5) / n_in
            # cells of ~REFINE_CELL_KEYS keys, but a band of at least -/+ 2 % of the window's keys: the
            # interpolation (linear in the DISTANCE, whose density is the smoother one) is good to ~1 %
            frac_cell = min(0.25, max(self.REFINE_CELL_KEYS / n_in, 0.04 / self.REFINE_CELLS))
            if any(abs(rel - x) < 0.25 * frac_cell for x in centres.get(c, [])):
                continue
            centres.setdefault(c, []).append(rel)
            half = self.REFINE_CELLS // 2
            d_lo, d_hi = np.sqrt(lo_v), np.sqrt(hi_v)
            for k in range(-half, half + 1):
                f = rel + k * frac_cell
                if 0.0 < f < 1.0:
                    dd = d_lo + (d_hi - d_lo) * f
                    t = to_bits(float(dd * dd))
                    if int(lows[c]) < t < int(thr[c]):
                        bands.setdefault(c, set()).add(t)
        if not bands:
            return thr, counts
        # one sweep per refined window over [window start, window end) only: the lag table then
        # resolves the narrow cells (nothing beyond the window is swept, pairs below it are bulk)
        out_thr, out_cnt = [], []
        for c in range(thr.size):
            if c not in bands:
                out_thr.append(int(thr[c]))
                out_cnt.append(int(counts[c]))
                continue
            inner = sorted(bands[c])
            t2 = np.array(([int(lows[c])] if int(lows[c]) > 0 else []) + inner + [int(thr[c])], dtype=np.uint64)
            with _phase("count sweep (%d classes, window)" % t2.size):
                c2 = self.hist_counts(t2)
            c2 = c2[1:] if int(lows[c]) > 0 else c2          # drop the class below the window
            if int(c2.sum()) != int(counts[c]):
                raise RuntimeError("window refinement count mismatch: %d != %d" % (int(c2.sum()), int(counts[c])))
            if _DEBUG_TIMES:
                cc = np.cumsum(c2)
                for r in ranks:
                    rr = r - (int(cum[c - 1]) if c else 0)
                    print("   [interp] rank at %.4f of the window, predicted cell %d of %d, found in cell %d; keys per cell ~%d, window keys %d"
                          % (rr / max(int(counts[c]), 1), len(inner) // 2, len(c2), int(np.searchsorted(cc, rr, side="right")),
                             int(np.median(c2[1:-1])) if len(c2) > 2 else 0, int(counts[c])))
            out_thr += inner + [int(thr[c])]
            out_cnt += [int(x) for x in c2]
        return np.array(out_thr, dtype=np.uint64), np.array(out_cnt, dtype=np.int64)

    def sq_order_stats(self, rank_fn, dirp=None, limit: Optional[float] = None):
        """Exact order statistics of the squared distances s of all (mask-passing)
        pairs with fl(sqrt(s)) <= limit.  rank_fn(total) -> 0-based ranks.
        Returns (dict rank -> s, total).

        Sweep 1: exact counts on a log-spaced grid of 4096 cells (sq_cell_counts).  A cell that
        holds a wanted rank spans 0.07 % of d: with few keys it is gathered directly (sweep 2);
        otherwise a fine bucket histogram and a gather run inside it (sweeps 2 and 3), where
        nearly every tile is culled.  Ties (crowded single values) refine further, exactly."""
        from ._exact import sq_upper_inclusive_bits
        ub = self.sq_upper_bound()
        hi_bits = to_bits(ub) + 1
        if limit is not None:
            hi_bits = min(hi_bits, sq_upper_inclusive_bits(limit))
        zero_lo = 0
        if self.pair_filter is not None:          # sparse space: 0 < d <= max_dist only
            hi_bits = min(hi_bits, self.pair_filter[0])
            zero_lo = 1 if self.pair_filter[1] else 0
        cells = None
        if (dirp is None and self.pair_filter is None and self.dim == 2
                and self.n_pairs >= self.SUBSAMPLE_SELECT_MIN_PAIRS and self._select_mode != "cells"):
            cells = self._window_cell_counts(rank_fn, hi_bits, limit)
        thr, counts = cells if cells is not None else self.sq_cell_counts(dirp, zero_lo, hi_bits)
        total = int(counts.sum())
        ranks = sorted(set(int(r) for r in rank_fn(total)))
        for r in ranks:
            if not (0 <= r < total):
                raise ValueError("rank %d outside [0, %d)" % (r, total))
        if not ranks:
            return {}, total
        cum = np.cumsum(counts)
        cls = np.searchsorted(cum, np.asarray(ranks, dtype=np.int64), side="right")
        wanted = sorted(set(int(c) for c in cls))
        # window table: [lo_1, hi_1, lo_2, hi_2, ...]; segment 2w+1 = inside window w
        bounds, win_lo, win_hi, base = [], [], [], []
        for c in wanted:
            lo = int(thr[c - 1]) if c > 0 else zero_lo
            hi = int(thr[c])
            bounds += [lo, hi]
            win_lo.append(lo)
            win_hi.append(hi)
            base.append(int(cum[c - 1]) if c > 0 else 0)
        wthr = np.array(bounds, dtype=np.uint64)
        nseg = len(bounds)
        lo_bits = [0] * nseg
        hi_bits_l = [0] * nseg
        for w in range(len(wanted)):
            lo_bits[2 * w + 1] = win_lo[w]
            hi_bits_l[2 * w + 1] = win_hi[w]
        seg_of_class = {c: 2 * w + 1 for w, c in enumerate(wanted)}
        seg_ranks = {}
        for r, c in zip(ranks, cls):
            seg_ranks.setdefault(seg_of_class[int(c)], []).append(r - base[wanted.index(int(c))])
        biggest = int(counts[wanted].max())
        direct = biggest <= self.DIRECT_GATHER_LIMIT and int(counts[wanted].sum()) <= (1 << 23)
        if direct:
            # the wanted cells are small: one bucket per cell, counts already known -> no
            # histogram sweep, straight to the gather
            nb = 1
            known = np.zeros((nseg, 1), dtype=np.int64)
            for w, c in enumerate(wanted):
                known[2 * w + 1, 0] = counts[c]
            _, rg = self._select_runner(_lib.KEY_SQDIST, wthr, dirp, nb)
            calls = []

            def rh(lo, hi, sc):
                if calls:   # a second histogram request would mean a refinement the known counts cannot serve
                    raise RuntimeError("direct gather asked for a refinement")
                calls.append(1)
                return known
        else:
            nb = 1 << 10
            while nb < (1 << 16) and nb * 256 < biggest:
                nb <<= 1
            while nseg * nb > (1 << 22) and nb > 256:
                nb >>= 1
            rh, rg = self._select_runner(_lib.KEY_SQDIST, wthr, dirp, nb)
        # direct mode: one bucket per cell whatever its size (never refine by count; a cell that is a
        # single bit pattern is still answered without a gather)
        vals, totals = exact_select(nseg, lo_bits, hi_bits_l, nb,
                                    lambda g, t: seg_ranks.get(g, []), rh, rg,
                                    gather_limit=1 << 24,
                                    bucket_limit=(1 << 62) if direct else (1 << 22))
        out = {}
        for r, c in zip(ranks, cls):
            w = wanted.index(int(c))
            seg = 2 * w + 1
            if totals[seg] != int(counts[int(c)]):
                raise RuntimeError("cell / window count mismatch in cell %d" % int(c))
            out[r] = vals[seg][r - base[w]]
        return out, total

    def sq_max(self, dirp=None):
        """Largest squared distance among the (mask-passing) pairs (None if no pair
        passes): one sweep with an empty histogram window, i.e. no atomics."""
        if self.pair_filter is not None:          # sparse space: the largest s inside the filter
            vals, total = self.sq_order_stats(lambda t: [t - 1] if t else [], dirp=dirp)
            return vals[total - 1] if total else None
        mx = torch.zeros(1, dtype=torch.int64, device=self.device)
        hist = torch.zeros(1, dtype=torch.int64, device=self.device)
        scratch = self._hist_scratch(1)
        zero = np.zeros(1, dtype=np.uint64)
        sc = np.zeros(1, dtype=np.float64)
        with torch.cuda.device(self.device):
            rc = self.lib.skg_pair_select_hist(
                _ptr(self.coords), _ptr(self.values), self.nv, self.vmode, self.n, self.dim, _hptr(dirp), None, 0,
                _lib.KEY_SQDIST, _hptr(zero), _hptr(zero), _hptr(sc), 1, _ptr(hist), _ptr(mx),
                _ptr(scratch), scratch.numel(), self.t0, self.t1, self.tstride, self._stream())
        _lib.check(rc, "skg_pair_select_hist")
        self.launches += 5 if self.t1 > self.t0 else 0
        amb = self.ambiguous_pairs(dirp)
        if amb is not None and amb[0].size:
            s_bits, _ = self._amb_keys(amb)
            s_bits = s_bits[s_bits <= INF_BITS]
            if s_bits.size:
                mx = torch.maximum(mx, torch.tensor([int(s_bits.max())], dtype=torch.int64,
                                                    device=self.device))
        D.allreduce_max_(mx, self.group)
        bits = int(mx.item())
        # bits == 0 means s == 0 for every pair or no pair at all; both give max d = 0 / none
        return from_bits(bits) if bits > 0 else (0.0 if (self.n_pairs > 0 and dirp is None) else None)

    WINDOWED_MEDIAN_MIN_PAIRS = 2 * 10 ** 8   # below this the plain two-sweep select is fast enough
    WINDOW_HALF_QUANTILE = 0.025   # ~7 sigma of the sub-sample's rank error; a miss falls back to the full range

    def _dz_select(self, thr, dirp, lo_bits, hi_bits, rank_fn, n_keys=None, first_hist=None, nb=None):
        """first_hist: the histogram of the first pass if a merged sweep already produced it (same
        windows, scales and bucket count as exact_select's first request)."""
        nl = int(thr.size)
        if nb is None:
            nb = self._buckets_for(self.n_pairs if n_keys is None else n_keys, nl)
        rh, rg = self._select_runner(_lib.KEY_ABSDZ, thr, dirp, nb)
        if first_hist is not None:
            calls = []

            def rh_first(lo, hi, sc, _rh=rh):
                if calls:
                    return _rh(lo, hi, sc)
                calls.append(1)
                return first_hist
            rh = rh_first
        return exact_select(nl, lo_bits, hi_bits, nb, rank_fn, rh, rg)

    def dz_windows(self, thr: np.ndarray, w_lo: np.ndarray, w_hi: np.ndarray, n_buckets: int, dirp=None):
        """One sweep (skg_pair_dz_windows): per lag class the exact pair count, the count of pairs with
        |dz| below the class's window [w_lo, w_hi) (bit patterns) and the bucket histogram of the
        |dz| inside the window - the first two passes of the windowed Dowd median in one."""
        from .select import _scale_for
        nl = int(thr.size)
        w_lo = np.ascontiguousarray(w_lo, dtype=np.uint64)
        w_hi = np.ascontiguousarray(w_hi, dtype=np.uint64)
        sc = np.array([_scale_for(int(a), int(b), n_buckets) if b > a else 0.0 for a, b in zip(w_lo, w_hi)])
        cnt = self._buffer("hist_cnt", nl, torch.int64).zero_()
        below = self._buffer("hist_below", nl, torch.float64).zero_()
        hist = self._buffer("select_hist", nl * n_buckets, torch.int64).zero_()
        scratch = self._hist_scratch(nl)
        with torch.cuda.device(self.device):
            rc = self.lib.skg_pair_dz_windows(
                _ptr(self.coords), _ptr(self.values), self.nv, self.vmode, self.n, self.dim, _hptr(dirp), _hptr(thr),
                nl, _hptr(w_lo), _hptr(w_hi), _hptr(sc), n_buckets, _ptr(cnt), _ptr(below), _ptr(hist),
                _ptr(scratch), scratch.numel(), self.t0, self.t1, self.tstride, self._stream())
        _lib.check(rc, "skg_pair_dz_windows")
        self.launches += 5 if self.t1 > self.t0 else 0
        both = torch.cat([cnt.to(torch.float64), below])
        amb = self.ambiguous_pairs(dirp)
        if amb is not None and amb[0].size:
            s_bits, dz = self._amb_keys(amb)
            g = np.searchsorted(thr, s_bits, side="right")
            keep = g < nl
            g, dz = g[keep], dz[keep]
            lo_v, kb = w_lo.view(np.float64), np.ascontiguousarray(dz).view(np.uint64)
            extra = np.zeros(2 * nl)
            np.add.at(extra, g, 1.0)
            np.add.at(extra, nl + g, (dz < lo_v[g]).astype(float))
            both += to_device(extra, self.device)
            inside = (kb >= w_lo[g]) & (kb < w_hi[g])
            t = (dz[inside] - lo_v[g[inside]]) * sc[g[inside]]
            b = np.minimum(n_buckets - 1, np.where(t >= 9.0e18, n_buckets - 1, t).astype(np.int64))
            hist += to_device(np.bincount(g[inside] * n_buckets + b, minlength=nl * n_buckets).astype(np.int64),
                              self.device)
        D.allreduce_sum_(both, self.group)
        D.allreduce_sum_(hist, self.group)
        host = to_host(both)
        return host[:nl].astype(np.int64), host[nl:].astype(np.int64), to_host(hist).reshape(nl, n_buckets), sc

    def hist_below(self, thr: np.ndarray, dz_lo: np.ndarray, dirp=None):
        """Exact per-lag (pair count, count of pairs with |dz| < dz_lo[lag]) with the fused-pass
        kernel (private counters, no atomics); all-reduced over the group."""
        thr, fsrc = self._filter_thr(np.asarray(thr, np.uint64))
        if fsrc is not None:   # per-lag bounds follow their classes; the s == 0 class gets 0
            lo2 = np.zeros(thr.size)
            ok = fsrc >= 0
            lo2[fsrc[ok]] = np.asarray(dz_lo, dtype=np.float64)[ok]
            dz_lo = lo2
        nl = int(thr.size)
        cnt = self._buffer("hist_cnt", nl, torch.int64).zero_()
        below = self._buffer("hist_below", nl, torch.float64).zero_()
        scratch = self._hist_scratch(nl)
        dz_lo = np.ascontiguousarray(dz_lo, dtype=np.float64)
        with torch.cuda.device(self.device):
            rc = self.lib.skg_pair_hist(
                _ptr(self.coords), _ptr(self.values), self.nv, self.vmode, self.n, self.dim, _hptr(dirp), _hptr(thr), nl,
                _lib.STAT_COUNT_BELOW, _hptr(dz_lo), _ptr(cnt), _ptr(below), None, _ptr(scratch),
                scratch.numel(), self.t0, self.t1, self.tstride, self._stream())
        _lib.check(rc, "skg_pair_hist")
        self.launches += 5 if self.t1 > self.t0 else 0
        both = torch.cat([cnt.to(torch.float64), below])
        amb = self.ambiguous_pairs(dirp)
        if amb is not None and amb[0].size:
            s_bits, dz = self._amb_keys(amb)
            g = np.searchsorted(thr, s_bits, side="right")
            keep = g < nl
            extra = np.zeros(2 * nl)
            np.add.at(extra, g[keep], 1.0)
            np.add.at(extra, nl + g[keep], (dz[keep] < dz_lo[g[keep]]).astype(float))
            both += to_device(extra, self.device)
        D.allreduce_sum_(both, self.group)
        host = to_host(both)
        return self._unmap(host[:nl].astype(np.int64), fsrc), self._unmap(host[nl:].astype(np.int64), fsrc)

    def dz_abs_sums(self, edges, dirp=None, thr_bits=None):
        """Per lag class (count, sum of the pair differences): fused pass over the SQUARED
        difference with the sum-of-square-roots statistic (sqrt(fl(x*x)) == |x|)."""
        if self.values is None:
            raise ValueError("values not set")
        thr = sq_threshold_bits(edges) if thr_bits is None else np.asarray(thr_bits, np.uint64)
        thr, fsrc = self._filter_thr(thr)
        nl = int(thr.size)
        out = self._buffer("hist_out", 3 * nl, torch.float64).zero_()
        cnt = self._buffer("hist_cnt", nl, torch.int64).zero_()
        scratch = self._hist_scratch(nl)
        with torch.cuda.device(self.device):
            rc = self.lib.skg_pair_hist(
                _ptr(self.coords), _ptr(self.values), self.nv, self.vmode | _lib.VALUE_SQUARE_FLAG,
                self.n, self.dim, _hptr(dirp), _hptr(thr), nl, _lib.STAT_SUM_SQRT, None, _ptr(cnt),
                None, _ptr(out[2 * nl:]), _ptr(scratch), scratch.numel(), self.t0, self.t1,
                self.tstride, self._stream())
        _lib.check(rc, "skg_pair_hist")
        self.launches += 5 if self.t1 > self.t0 else 0
        out[:nl] = cnt.to(torch.float64)
        amb = self.ambiguous_pairs(dirp)
        if amb is not None and amb[0].size:
            s_bits, dz = self._amb_keys(amb)
            g = np.searchsorted(thr, s_bits, side="right")
            keep = g < nl
            extra = np.zeros(3 * nl)
            np.add.at(extra, g[keep], 1.0)
            np.add.at(extra, 2 * nl + g[keep], dz[keep])
            out += to_device(extra, self.device)
        D.allreduce_sum_(out, self.group)
        host = to_host(out)
        return self._unmap(host[:nl].astype(np.int64), fsrc), self._unmap(host[2 * nl:].copy(), fsrc)

    def dz_order_stats(self, edges, rank_fn, dirp=None, thr_bits=None):
        """Exact order statistics of the pair differences per lag class: rank_fn(lag, total) ->
        0-based ranks.  Returns (list of dict rank -> value, totals)."""
        if self.values is None:
            raise ValueError("values not set")
        thr = sq_threshold_bits(edges) if thr_bits is None else np.asarray(thr_bits, np.uint64)
        thr, fsrc = self._filter_thr(thr)
        nl = int(thr.size)
        hi = to_bits(self._dz_bound) + 1
        if fsrc is None:
            return self._dz_select(thr, dirp, [0] * nl, [hi] * nl, rank_fn)
        back = {int(c): k for k, c in enumerate(fsrc) if c >= 0}   # filtered class -> user class
        vals2, tot2 = self._dz_select(thr, dirp, [0] * nl, [hi] * nl,
                                      lambda g, t: rank_fn(back[g], t) if g in back else [])
        vals = [vals2[c] if c >= 0 else {} for c in fsrc]
        tots = [tot2[c] if c >= 0 else 0 for c in fsrc]
        return vals, tots

    def dz_value_histogram(self, edges, value_edges, dirp=None, thr_bits=None):
        """np.histogram(|dz| of lag class g, bins=value_edges)[0] for every lag class g: exact
        counts [n_lags, len(value_edges) - 1] from one below-count sweep per value edge (numpy's
        bins are half open, the last one closed)."""
        thr = sq_threshold_bits(edges) if thr_bits is None else np.asarray(thr_bits, np.uint64)
        nl = int(thr.size)
        ve = np.asarray(value_edges, dtype=np.float64)
        if ve.ndim != 1 or ve.size < 2 or np.any(ve[:-1] > ve[1:]):
            raise ValueError("value_edges must be an increasing 1-D array")
        cuts = ve.copy()
        cuts[-1] = np.nextafter(ve[-1], np.inf)      # x <= last edge  <=>  x < nextafter(last edge)
        below = np.empty((cuts.size, nl), dtype=np.int64)
        for k, cut in enumerate(cuts):
            _, below[k] = self.hist_below(thr, np.full(nl, cut), dirp)
        return np.diff(below, axis=0).T

    def dz_medians(self, edges, dirp=None, thr_bits=None):
        """Per lag class: (count, median |dz|) exactly as np.nanmedian would return
        over the materialised lag class (estimators.py:178).

        Small problems: histogram of every in-range |dz| + gather (two sweeps, one L2 atomic
        per in-range pair).  Large problems: (1) per-lag windows around the median from a
        random sub-sample of the points; (2) exact per-lag counts and below-window counts with
        the fused-pass kernel (no atomics); (3)+(4) fine histogram / gather inside the windows
        only (~8 % of the in-range pairs touch an atomic).  A lag whose window misses the
        median falls back to the full range, so the result is always exact."""
        if self.values is None:
            raise ValueError("values not set")
        thr = sq_threshold_bits(edges) if thr_bits is None else np.asarray(thr_bits, np.uint64)
        thr, fsrc = self._filter_thr(thr)
        nl = int(thr.size)
        hi = to_bits(self._dz_bound) + 1  # >= every pair's difference

        def rank_fn(g, tot):
            return [] if tot == 0 else [(tot - 1) // 2, tot // 2]

        totals = med_lo = med_hi = None
        # NaN observations: np.nanmedian ranks only the finite differences, which the windowed
        # path's pair counts cannot tell apart -> plain select over the finite keys
        finite_values = bool(np.isfinite(self._orig_values).all())
        if finite_values and self.n_pairs >= self.WINDOWED_MEDIAN_MIN_PAIRS and nl * 16 * 64 <= 180 * 1024:
            win = self._median_windows(thr, dirp, hi)
            if win is not None:
                w_lo, w_hi = self._merge_dz_windows(*win, hi)   # per-lag bit windows [lo, hi)
                if fsrc is None:
                    # one sweep: counts, below-window counts and the in-window histogram
                    nbk = 1 << 12
                    with _phase("dz windows sweep"):
                        counts, below, hist0, _ = self.dz_windows(thr, w_lo, w_hi, nbk, dirp)
                else:
                    dz_lo = w_lo.view(np.float64)
                    counts, below = self.hist_below(thr, dz_lo, dirp)
                    hist0, nbk = None, None
                ra = (counts - 1) // 2 - below       # ranks relative to the window start
                rb = counts // 2 - below

                def rel_rank_fn(g, tot):
                    if counts[g] == 0 or ra[g] < 0 or rb[g] >= tot:
                        return []                     # window missed the median: handled below
                    return [int(ra[g]), int(rb[g])]

                # the windows hold ~2 * WINDOW_HALF_QUANTILE of the in-range pairs: size the bucket
                # histogram (device memory, D2H copy, host prefix sums) for that population
                vals, wtot = self._dz_select(thr, dirp, [int(x) for x in w_lo],
                                             [int(x) for x in w_hi], rel_rank_fn,
                                             n_keys=int(counts.sum() * 2.5 * self.WINDOW_HALF_QUANTILE) + 1,
                                             first_hist=hist0, nb=nbk)
                totals = [int(c) for c in counts]
                med_lo = [None] * nl
                med_hi = [None] * nl
                missed = []
                for g in range(nl):
                    if counts[g] == 0:
                        continue
                    if ra[g] >= 0 and rb[g] < wtot[g]:
                        med_lo[g], med_hi[g] = vals[g][int(ra[g])], vals[g][int(rb[g])]
                    else:
                        missed.append(g)
                if missed:  # exact fallback on the full range for those lags only
                    lo_b = [0] * nl
                    hi_b = [hi if g in missed else 0 for g in range(nl)]
                    v2, t2 = self._dz_select(thr, dirp, lo_b, hi_b, rank_fn)
                    for g in missed:
                        med_lo[g], med_hi[g] = v2[g][(t2[g] - 1) // 2], v2[g][t2[g] // 2]
        if totals is None:
            vals, totals = self._dz_select(thr, dirp, [0] * nl, [hi] * nl, rank_fn)
            med_lo = [vals[g].get((totals[g] - 1) // 2) if totals[g] else None for g in range(nl)]
            med_hi = [vals[g].get(totals[g] // 2) if totals[g] else None for g in range(nl)]
        med = np.full(nl, np.nan)
        for g in range(nl):
            t = totals[g]
            if t:
                a, b = med_lo[g], med_hi[g]
                # np.median: mean of the two middle elements (numpy _median -> mean)
                med[g] = a if (t % 2) else (a + b) / 2.0
        return self._unmap(np.asarray(totals, dtype=np.int64), fsrc), self._unmap(med, fsrc, fill=np.nan)

    WINDOW_MERGE_GROWTH = 2.0

    @classmethod
    def _merge_dz_windows(cls, w_lo, w_hi, hi_bits_full):
        """Runs of consecutive lag classes whose windows nearly coincide (the classes past the
        range of the field share one median) get ONE common window, their union, as long as it
        is at most WINDOW_MERGE_GROWTH times as wide as the widest member: the sweep then needs
        no per-class window selection for pairs whose candidate classes lie in one run
        (pair_win.cu, STATS == 4).  Any window that holds the median is a valid window."""
        nl = int(w_lo.size)
        lo = w_lo.copy()
        hi = w_hi.copy()
        lo_f, hi_f = w_lo.view(np.float64), w_hi.view(np.float64)
        g = 0
        while g < nl:
            h = g + 1
            if not (w_lo[g] == 0 and w_hi[g] >= hi_bits_full):      # full-range windows stay alone
                a, b = int(w_lo[g]), int(w_hi[g])
                width = hi_f[g] - lo_f[g]
                while h < nl and not (w_lo[h] == 0 and w_hi[h] >= hi_bits_full):
                    na, nb = min(a, int(w_lo[h])), max(b, int(w_hi[h]))
                    wmax = max(width, hi_f[h] - lo_f[h])
                    if from_bits(nb) - from_bits(na) > cls.WINDOW_MERGE_GROWTH * wmax:
                        break
                    a, b, width = na, nb, wmax
                    h += 1
                lo[g:h] = a
                hi[g:h] = b
            g = h
        return lo, hi

    def _median_windows(self, thr, dirp, hi_bits_full):
        """Per-lag |dz| windows (bit patterns) that very probably hold the lag's median: the
        0.5 -/+ WINDOW_HALF_QUANTILE quantiles of a random sub-sample of the points.  The same
        sample (fixed seed) on every rank, so all ranks use identical windows."""
        nl = int(thr.size)
        if self.n <= self.SUBSAMPLE_POINTS:
            return None
        sub = self._subsample_engine()
        if getattr(sub, "_parent_values_version", None) != self._values_version or sub.vmode != self.vmode:
            sub.set_values(self._orig_values[self._sub_idx], self.vmode)
            sub._parent_values_version = self._values_version
        sub.pair_filter = self.pair_filter
        q = self.WINDOW_HALF_QUANTILE

        def rank_fn(g, tot):
            if tot < 400:
                return []
            return [max(0, int((0.5 - q) * tot)), min(tot - 1, int((0.5 + q) * tot))]

        with _phase("dz sub-sample select"):
            vals, totals = sub._dz_select(thr, dirp, [0] * nl, [hi_bits_full] * nl, rank_fn)
        self.launches += sub.launches
        sub.launches = 0
        sub.pair_filter = None
        w_lo = np.zeros(nl, dtype=np.uint64)
        w_hi = np.full(nl, hi_bits_full, dtype=np.uint64)
        for g in range(nl):
            r = rank_fn(g, totals[g])
            if r:
                w_lo[g] = to_bits(vals[g][r[0]])
                w_hi[g] = to_bits(vals[g][r[1]]) + 1
        return w_lo, w_hi

    # -------------------------------------------------------- materialisation
    def materialize(self, want=("dist",), edges=None, dirp=None, chunk=1 << 24, _raw=False):
        """Condensed arrays (Variogram.distance / pairwise_diffs / lag_groups /
        direction mask) produced on the device in chunks and copied to the host."""
        thr = None if edges is None else sq_threshold_bits(edges)
        nl = 0 if thr is None else int(thr.size)
        P = self.n_pairs
        out = {}
        if "dist" in want:
            out["dist"] = np.empty(P, dtype=np.float64)
        if "diffs" in want:
            out["diffs"] = np.empty(P, dtype=np.float64)
        if "groups" in want:
            out["groups"] = np.empty(P, dtype=np.int64)
        if "mask" in want:
            out["mask"] = np.empty(P, dtype=np.uint8)
        for k0 in range(0, P, chunk):
            k1 = min(P, k0 + chunk)
            m = k1 - k0
            bufs = {}
            if "dist" in want:
                bufs["dist"] = torch.empty(m, dtype=torch.float64, device=self.device)
            if "diffs" in want:
                bufs["diffs"] = torch.empty(m, dtype=torch.float64, device=self.device)
            if "groups" in want:
                bufs["groups"] = torch.empty(m, dtype=torch.int64, device=self.device)
            if "mask" in want:
                bufs["mask"] = torch.empty(m, dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                rc = self.lib.skg_pair_materialize(
                    _ptr(self._coords_unsorted), _ptr(self._values_unsorted), self.nv, self.vmode, self.n, self.dim,
                    _hptr(dirp),
                    _hptr(thr), nl, _ptr(bufs.get("dist")), _ptr(bufs.get("diffs")),
                    _ptr(bufs.get("groups")), _ptr(bufs.get("mask")), k0, k1, self._stream())
            _lib.check(rc, "skg_pair_materialize")
            self.launches += 1
            for name, b in bufs.items():
                out[name][k0:k1] = to_host(b)
        amb = self.ambiguous_pairs(dirp, all_tiles=True)
        if amb is not None and amb[0].size and ("mask" in want or "groups" in want):
            i, j = self.perm[amb[0]], self.perm[amb[1]]  # caller's order, i < j by construction
            k = self.n * i - i * (i + 1) // 2 + (j - i - 1)
            if "mask" in want:
                out["mask"][k] = 1
            if "groups" in want:
                s_bits, _ = self._amb_keys(amb)
                g = np.searchsorted(thr, s_bits, side="right")
                out["groups"][k] = np.where(g >= nl, -1, g)
        if self.pair_filter is not None and not _raw:
            out = self._sparse_view(out, want)
        return out

    def _sparse_view(self, out, want):
        """Sparse (max_dist) spaces: keep the pairs with 0 < d <= max_dist and put them in the
        order of the reference's triangular CSR matrix (Variogram.py:1210-1227: lower-triangle
        entries row by row, i.e. pairs (i < j) sorted by j, then i)."""
        n = self.n
        if "dist" in out:
            d = out["dist"]
        else:
            d = self.materialize(("dist",), _raw=True)["dist"]
        keep = np.flatnonzero((d > 0) & (d <= self.max_dist))
        iu, ju = np.triu_indices(n, 1)
        order = keep[np.lexsort((iu[keep], ju[keep]))]
        return {name: out[name][order] for name in want}


class KrigeEngine:
    """Batched ordinary kriging over one sample set resident in HBM."""

    CELL_TARGET = 4.0  # samples per grid cell aimed for

    def __init__(self, coords, values, radius: float, device=None, group=None):
        _require_cuda()
        self.lib = _lib.load()
        c = np.ascontiguousarray(np.asarray(coords, dtype=np.float64))
        if c.ndim != 2 or c.shape[1] not in (2, 3):
            raise ValueError("coords must be (N, 2) or (N, 3); got %r" % (c.shape,))
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64)).ravel()
        if v.size != c.shape[0]:
            raise ValueError("values length does not match coordinates")
        self.n, self.dim = int(c.shape[0]), int(c.shape[1])
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.group = group
        self.rank, self.world = D.world(group)
        self.radius = float(radius)
        self.launches = 0
        self.grid = self._grid_params(c, self.radius)
        ncell = int(self.grid[4] * self.grid[5] * self.grid[6])
        with torch.cuda.device(self.device):
            self.coords = to_device(c.T, self.device)
            self.values = to_device(v, self.device)
            self.cell_start = torch.empty(ncell + 1, dtype=torch.int32, device=self.device)
            self.order = torch.empty(self.n, dtype=torch.int32, device=self.device)
            scratch = torch.empty((ncell + self.n) * 4, dtype=torch.uint8, device=self.device)
            rc = self.lib.skg_krige_build_grid(
                _ptr(self.coords), self.n, self.dim, _hptr(self.grid), _ptr(self.cell_start),
                _ptr(self.order), _ptr(scratch), scratch.numel(),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _lib.check(rc, "skg_krige_build_grid")
        self.launches += 4
        self._scratch = torch.empty(16, dtype=torch.uint8, device=self.device)
        self._cov, self._cov_key = None, None

    SAMPLE_COV_MAX_BYTES = 4 << 30   # cache the N x N sample covariance up to this size (N <= 23170)

    def _sample_cov(self, model, mp, matrix_lut):
        """c_t - gamma(d_ij) of all sample pairs on the device (skg_krige_sample_cov), built once per
        (model, parameters) and reused by every transform: the per-target matrix build becomes a
        gather.  None when N x N doubles exceed SAMPLE_COV_MAX_BYTES (computed on the fly then)."""
        if self.n * self.n * 8 > self.SAMPLE_COV_MAX_BYTES:
            return None
        key = (model, tuple(float(x) for x in mp), None if matrix_lut is None else int(matrix_lut.data_ptr()))
        if self._cov_key != key:
            if self._cov is None:
                self._cov = torch.empty(self.n * self.n, dtype=torch.float64, device=self.device)
            rc = self.lib.skg_krige_sample_cov(
                _ptr(self.coords), self.n, self.dim, _lib.MODEL_IDS[model], _hptr(np.ascontiguousarray(mp)),
                _ptr(matrix_lut), 0 if matrix_lut is None else int(matrix_lut.numel()), _ptr(self._cov),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            _lib.check(rc, "skg_krige_sample_cov")
            self.launches += 1
            self._cov_key = key
        return self._cov

    @classmethod
    def _grid_params(cls, c, radius):
        lo, hi = c.min(axis=0), c.max(axis=0)
        ext = hi - lo
        pos = ext[ext > 0]
        if pos.size:
            h = float((np.prod(pos) * cls.CELL_TARGET / len(c)) ** (1.0 / pos.size))
        else:
            h = 1.0
        if radius > 0 and np.isfinite(radius):
            h = max(h, radius / 64.0)
        h = max(h, float(ext.max()) / 2048.0, 1e-300)
        dims = [int(np.floor(e / h)) + 1 for e in ext]
        while np.prod([float(d) for d in dims]) > float(1 << 24):
            h *= 1.5
            dims = [int(np.floor(e / h)) + 1 for e in ext]
        g = np.zeros(8, dtype=np.float64)
        g[: len(lo)] = lo
        g[3] = h
        g[4], g[5] = dims[0], dims[1]
        g[6] = dims[2] if len(dims) > 2 else 1
        return g

    def transform(self, targets, model: str, model_params, min_points: int, max_points: int,
                  exclude=None, matrix_lut=None, soa=False):
        """targets (M, dim) - or, with soa=True, (dim, M) as the caller's per-dimension arrays
        already are - -> (z, sigma, status) numpy arrays for ALL M targets; with a process group
        each rank krigs its contiguous slice and the slices are gathered."""
        t = np.asarray(targets, dtype=np.float64)
        if soa:
            if t.ndim != 2 or t.shape[0] != self.dim:
                raise ValueError("targets must be (%d, M)" % self.dim)
            t_soa = np.ascontiguousarray(t)
        else:
            if t.ndim != 2 or t.shape[1] != self.dim:
                raise ValueError("targets must be (M, %d)" % self.dim)
            t_soa = np.ascontiguousarray(t.T)
        m = int(t_soa.shape[1])
        mp = np.asarray(model_params, dtype=np.float64)
        if mp.size == 4:
            mp = np.concatenate((mp, [mp[0], mp[1]]))
        if mp.size != 6:
            raise ValueError("model_params = [range, sill, shape, nugget(, table range, table sill)]")
        with torch.cuda.device(self.device):
            tg = to_device(t_soa, self.device)
            ex = None
            if exclude is not None:
                ex = to_device(np.ascontiguousarray(exclude, dtype=np.int32), self.device)
            lut = None
            if matrix_lut is not None:
                lut = to_device(np.ascontiguousarray(matrix_lut, dtype=np.float64), self.device)
            z, sigma, status = self.transform_device(tg, m, model, mp, min_points, max_points, ex, lut)
            if self.world > 1:
                b, e = D.shard_range(m, self.rank, self.world)
                z = D.allgather_varlen(z[b:e], self.group)
                sigma = D.allgather_varlen(sigma[b:e], self.group)
                status = D.allgather_varlen(status[b:e], self.group)
        return to_host(z), to_host(sigma), to_host(status)

    def transform_device(self, targets_soa: torch.Tensor, m: int, model: str, mp: np.ndarray,
                         min_points: int, max_points: int, exclude: Optional[torch.Tensor] = None,
                         matrix_lut: Optional[torch.Tensor] = None):
        """Device-resident variant: targets_soa is a (dim*M) fp64 CUDA tensor; returns
        device tensors (only this rank's slice [begin, end) is filled)."""
        z = torch.full((m,), float("nan"), dtype=torch.float64, device=self.device)
        sigma = torch.full((m,), float("nan"), dtype=torch.float64, device=self.device)
        status = torch.zeros(m, dtype=torch.int32, device=self.device)
        b, e = D.shard_range(m, self.rank, self.world)
        mp = np.ascontiguousarray(mp, dtype=np.float64)
        if mp.size == 4:
            mp = np.concatenate((mp, [mp[0], mp[1]]))
        rc = self.lib.skg_krige(
            _ptr(self.coords), _ptr(self.values), self.n, self.dim, _hptr(self.grid),
            _ptr(self.cell_start), _ptr(self.order), _ptr(targets_soa), m,
            _lib.MODEL_IDS[model], _hptr(mp), _ptr(matrix_lut),
            0 if matrix_lut is None else int(matrix_lut.numel()), self.radius, int(min_points), int(max_points),
            _ptr(exclude), _ptr(self._sample_cov(model, mp, matrix_lut)), _ptr(z), _ptr(sigma), _ptr(status),
            _ptr(self._scratch), self._scratch.numel(),
            b, e, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _lib.check(rc, "skg_krige")
        self.launches += 2 if e > b else 0
        return z, sigma, status
