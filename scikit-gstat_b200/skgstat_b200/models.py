"""Theoretical variogram models gamma(h) on the host (numpy, vectorised).

Used by the CPU model fit (scipy curve_fit, which stays on the CPU) and by
``Variogram.fitted_model``.  Same closed forms and argument order
``(h, r, c0, [s], b=0)`` as the reference (models.py:21-422); the kriging kernel
holds a device restatement of the first five (csrc/krige_kernels.cu).
"""
from __future__ import annotations

import numpy as np


def _prep(h):
    arr = np.asarray(h, dtype=float)
    return arr, arr.ndim == 0


def _out(v, scalar):
    return float(v) if scalar else v


def spherical(h, r, c0, b=0.0):
    """models.py:77-82."""
    x, sc = _prep(h)
    q = x / (r / 1.0)
    v = np.where(x <= r, b + c0 * ((1.5 * q) - (0.5 * (q ** 3.0))), b + c0)
    return _out(v, sc)


def exponential(h, r, c0, b=0.0):
    """models.py:142-144."""
    x, sc = _prep(h)
    return _out(b + c0 * (1.0 - np.exp(-(x / (r / 3.0)))), sc)


def gaussian(h, r, c0, b=0.0):
    """models.py:207-209."""
    x, sc = _prep(h)
    a = r / 2.0
    return _out(b + c0 * (1.0 - np.exp(-(x ** 2 / a ** 2))), sc)


def cubic(h, r, c0, b=0.0):
    """models.py:264-272."""
    x, sc = _prep(h)
    a = r / 1.0
    inner = ((7 * (x ** 2 / a ** 2)) - ((35 / 4) * (x ** 3 / a ** 3))
             + ((7 / 2) * (x ** 5 / a ** 5)) - ((3 / 4) * (x ** 7 / a ** 7)))
    return _out(np.where(x < r, b + c0 * inner, b + c0), sc)


def stable(h, r, c0, s, b=0.0):
    """models.py:336-344."""
    x, sc = _prep(h)
    a = r / np.power(3, 1 / s)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = b + c0 * (1.0 - np.exp(-np.power(x / a, s)))
    return _out(np.where(x == 0, b, v), sc)


def matern(h, r, c0, s, b=0.0):
    """models.py:412-422 (host only; needs the Bessel function K_s)."""
    from scipy import special
    x, sc = _prep(h)
    a = r / 2.0
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        t = (x * np.sqrt(s)) / a
        v = b + c0 * (1.0 - (2 / special.gamma(s)) * np.power(t, s) * special.kv(s, 2 * t))
    return _out(np.where(x == 0, b, v), sc)


MODELS = dict(spherical=spherical, exponential=exponential, gaussian=gaussian, cubic=cubic,
              stable=stable, matern=matern)
# models the kriging kernel implements on the device
DEVICE_MODELS = ("spherical", "exponential", "gaussian", "cubic", "stable", "matern")
