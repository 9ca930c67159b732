"""CPU: the C-ABI library loads, exports every symbol include/skg_b200.h declares, and
rejects bad arguments without touching a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from skgstat_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "skg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(skg_[a-z0-9_]+)\s*\(", text)))


def test_library_present_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "build with make -C scikit-gstat_b200/csrc"
    lib = _lib.load()
    assert lib.skg_abi_version() == _lib.ABI_VERSION


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), "symbol %s missing from libskg_b200.so" % name
        assert name in _lib.SIGNATURES, "%s has no ctypes signature in _lib.py" % name
    assert sorted(_lib.SIGNATURES) == names


def test_tile_count_and_constants():
    assert _lib.num_tiles(0) == 0 and _lib.num_tiles(1) == 0
    assert _lib.num_tiles(2) == 1
    assert _lib.num_tiles(256) == 1
    assert _lib.num_tiles(257) == 3
    assert _lib.num_tiles(20000) == 79 * 80 // 2
    text = open(os.path.join(ROOT, "include", "skg_b200.h")).read()
    assert "#define SKG_TILE %d" % _lib.TILE in text
    assert "#define SKG_MAX_LAGS %d" % _lib.MAX_LAGS in text
    assert "#define SKG_KRIGE_MAX_POINTS %d" % _lib.KRIGE_MAX_POINTS in text


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    thr = np.array([1, 2, 3], dtype=np.uint64)
    rc = lib.skg_pair_hist(None, None, 1, 0, 10, 2, None, thr.ctypes.data_as(C.c_void_p), 3, 1,
                           None, None, None, None, None, 0, 0, 0, 1, None)
    assert rc == -1 and b"NULL" in lib.skg_last_error()
    rc = lib.skg_pair_materialize(C.c_void_p(8), None, 1, 0, 10, 5, None, None, 0, None, None, None,
                                  None, 0, 0, None)
    assert rc == -2 and b"dim" in lib.skg_last_error()
    rc = lib.skg_krige(None, None, 0, 2, None, None, None, None, 0, 0, None, None, 0, 1.0, 1, 1, None, None,
                       None, None, None, None, 0, 0, 0, None)
    assert rc == -1
    with pytest.raises(_lib.SkgError):
        _lib.check(rc, "skg_krige")
    # the newer entry points validate before touching the device too
    rc = lib.skg_pair_hist_batch(None, None, 4, 10, 2, None, thr.ctypes.data_as(C.c_void_p), 3, 1,
                                 None, None, None, 0, 0, 0, 1, None)
    assert rc == -1 and b"NULL" in lib.skg_last_error()
    rc = lib.skg_pair_cell_counts(None, 10, 2, None, 0, 1 << 62, 11, 0, None, None, 0, 0, 0, 1, None)
    assert rc == -1
    rc = lib.skg_pair_hist(C.c_void_p(8), C.c_void_p(8), 3, 1, 10, 2, None, thr.ctypes.data_as(C.c_void_p),
                           3, 1, None, C.c_void_p(8), C.c_void_p(8), None, None, 0, 0, 0, 1, None)
    assert rc == -1 and b"cross 2" in lib.skg_last_error()   # "scalar needs 1 column, cross 2, ..."
    assert lib.skg_pair_batch_max_lags() >= 64
    rc = lib.skg_pick_order_stats(None, None, None, 0, None, None, None, None, None, 1, None, None, 0, None)
    assert rc == -1 and b"skg_pick_order_stats" in lib.skg_last_error()
    off = np.zeros(40, dtype=np.int32)
    rc = lib.skg_pick_order_stats(C.c_void_p(8), C.c_void_p(8), C.c_void_p(8), 4, C.c_void_p(8),
                                  off.ctypes.data_as(C.c_void_p), C.c_void_p(8), None, None, 33, C.c_void_p(8),
                                  None, 0, None)
    assert rc == -1 and b"n_want" in lib.skg_last_error()
    wslot, wrank = np.zeros(4, dtype=np.uint32), np.zeros(4, dtype=np.int64)
    rc = lib.skg_pick_order_stats(C.c_void_p(8), C.c_void_p(8), C.c_void_p(8), 1 << 20,
                                  wslot.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                  wrank.ctypes.data_as(C.c_void_p), None, None,
                                  1, C.c_void_p(8), None, 0, None)
    assert rc == -1 and b"scratch" in lib.skg_last_error()
    assert lib.skg_pick_scratch_bytes() > 0


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from skgstat_b200.engine import PairEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PairEngine(np.zeros((4, 2)))
    import skgstat_b200 as skg
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        skg.Variogram(np.random.rand(10, 2), np.random.rand(10))


def test_makefile_builds_every_cuda_source():
    """build() runs `make -C scikit-gstat_b200/csrc`: a kernel file missing from SRC would compile
    nowhere and only show up as an undefined symbol on the GPU box."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "scikit-gstat_b200", "csrc")
    text = open(os.path.join(csrc, "Makefile")).read()
    listed = set(re.search(r"^SRC\s*:=\s*(.+)$", text, re.M).group(1).split())
    present = {os.path.basename(p) for p in glob.glob(os.path.join(csrc, "*.cu"))}
    assert listed == present
