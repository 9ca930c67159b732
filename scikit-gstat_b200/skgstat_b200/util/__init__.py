"""Callers of the hot path rebuilt on the GPU kernels."""
from . import cross_validation, cross_variogram, uncertainty  # noqa: F401
