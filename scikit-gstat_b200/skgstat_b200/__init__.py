"""skgstat_b200 - B200-native (sm_100a) experimental-variogram and
ordinary-kriging hot path behind scikit-gstat's API.

Same class names and constructor arguments as the reference (skgstat/__init__.py:1-9)
for the path named in BASELINE.json: Variogram, DirectionalVariogram,
OrdinaryKriging, MetricSpace.  Model fitting stays on the CPU (scipy curve_fit),
everything O(N^2) or O(M * max_points^3) runs in libskg_b200.so.
"""
from . import _lib  # noqa: F401  (does not load the .so until first use)
from .MetricSpace import MetricSpace, MetricSpacePair
from .Variogram import Variogram
from .DirectionalVariogram import DirectionalVariogram
from .Kriging import OrdinaryKriging
from . import binning, estimators, models, util  # noqa: F401

__version__ = "0.1.0"
__all__ = ["Variogram", "DirectionalVariogram", "OrdinaryKriging", "MetricSpace",
           "MetricSpacePair"]
