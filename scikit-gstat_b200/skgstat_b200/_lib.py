"""ctypes binding of the C ABI declared in include/skg_b200.h.

The shared library is built in-tree (``scikit-gstat_b200/csrc/Makefile`` or
``__graft_entry__.build()``) next to this file.  There is NO fallback: if the
library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SKG_B200_LIB") or os.path.join(_HERE, "libskg_b200.so")  # env: developer builds

ABI_VERSION = 6
TILE = 256
MAX_LAGS = 250
KRIGE_MAX_POINTS = 128
PICK_DIRECT_MAX = 16384

DIR_NONE, DIR_TRIANGLE, DIR_COMPASS = 0, 1, 2
STAT_SUM_SQ, STAT_SUM_SQRT, STAT_COUNT_BELOW = 1, 2, 4
KEY_SQDIST, KEY_ABSDZ = 0, 1
VALUE_SCALAR, VALUE_CROSS, VALUE_AGGREGATE = 0, 1, 2
VALUE_DIST_S, VALUE_DIST_D, VALUE_DIST_D15 = 3, 4, 5
VALUE_SQUARE_FLAG = 0x100
MAX_VALUE_COLUMNS = 8
CELL_COUNT = 4096
MODEL_IDS = dict(spherical=0, exponential=1, gaussian=2, cubic=3, stable=4, matern=5)
KRIGE_OK, KRIGE_LESS_POINTS, KRIGE_SINGULAR = 0, 1, 2

_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_dbl = C.c_double

# name -> (restype, argtypes); mirrors include/skg_b200.h one to one
SIGNATURES = {
    "skg_abi_version": (_int, []),
    "skg_last_error": (C.c_char_p, []),
    "skg_device_sm_count": (_int, [C.POINTER(_int)]),
    "skg_fp64_peak": (_int, [_int, C.POINTER(_dbl), C.POINTER(_dbl)]),
    "skg_morton_keys": (_int, [_p, _i64, _int, _p, _p, _p, _p]),
    "skg_pair_num_tiles": (_i64, [_i64]),
    "skg_pair_dir_ambiguous": (_int, [_p, _i64, _int, _p, _p, _p, _p, _i64, _p, _i64, _i64, _i64, _i64, _p]),
    "skg_pair_scratch_bytes": (_i64, [_i64, _int]),
    "skg_pair_hist": (_int, [_p, _p, _int, _int, _i64, _int, _p, _p, _int, _int, _p, _p, _p, _p, _p, _i64,
                             _i64, _i64, _i64, _p]),
    "skg_pair_dz_windows": (_int, [_p, _p, _int, _int, _i64, _int, _p, _p, _int, _p, _p, _p, _i64, _p, _p, _p, _p,
                                   _i64, _i64, _i64, _i64, _p]),
    "skg_pair_batch_max_lags": (_int, []),
    "skg_pair_batch_scratch_bytes": (_i64, [_i64, _int]),
    "skg_pair_hist_batch": (_int, [_p, _p, _int, _i64, _int, _p, _p, _int, _int, _p, _p, _p, _i64,
                                   _i64, _i64, _i64, _p]),
    "skg_pair_cell_scratch_bytes": (_i64, [_i64]),
    "skg_pair_cell_counts": (_int, [_p, _i64, _int, _p, C.c_uint64, C.c_uint64, _int, _int, _p, _p, _i64,
                                    _i64, _i64, _i64, _p]),
    "skg_pair_select_hist": (_int, [_p, _p, _int, _int, _i64, _int, _p, _p, _int, _int, _p, _p, _p, _i64,
                                    _p, _p, _p, _i64, _i64, _i64, _i64, _p]),
    "skg_pair_select_gather": (_int, [_p, _p, _int, _int, _i64, _int, _p, _p, _int, _int, _p, _p, _p, _i64,
                                      _p, _p, _p, _p, _i64, _p, _i64, _i64, _i64, _i64, _p]),
    "skg_pick_scratch_bytes": (_i64, []),
    "skg_pick_order_stats": (_int, [_p, _p, _p, _i64, _p, _p, _p, _p, _p, _int, _p, _p, _i64, _p]),
    "skg_pair_materialize": (_int, [_p, _p, _int, _int, _i64, _int, _p, _p, _int, _p, _p, _p, _p,
                                    _i64, _i64, _p]),
    "skg_krige_build_grid": (_int, [_p, _i64, _int, _p, _p, _p, _p, _i64, _p]),
    "skg_krige_scratch_bytes": (_i64, [_int]),
    "skg_krige_sample_cov": (_int, [_p, _i64, _int, _int, _p, _p, _int, _p, _p]),
    "skg_krige": (_int, [_p, _p, _i64, _int, _p, _p, _p, _p, _i64, _int, _p, _p, _int, _dbl, _int, _int,
                         _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _p]),
}

_lib = None


class SkgError(RuntimeError):
    """A C-ABI call returned non-zero."""


def load():
    """Load libskg_b200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libskg_b200.so not found at %s - build it with `make -C scikit-gstat_b200/csrc` "
            "or `python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU "
            "fallback for the hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.skg_abi_version() != ABI_VERSION:
        raise ImportError("libskg_b200.so ABI %d != expected %d"
                          % (lib.skg_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().skg_last_error().decode("utf-8", "replace")
        raise SkgError("%s failed (rc=%d): %s" % (what, rc, msg))


def num_tiles(n: int) -> int:
    return int(load().skg_pair_num_tiles(int(n)))
