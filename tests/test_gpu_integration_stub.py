"""The binding INTEGRATION.md shows a scikit-gstat maintainer (raw ctypes on libskg_b200.so, torch
only for device buffers - no skgstat_b200.engine in between) against the oracle: proves the
documented C-ABI call sequence is complete and current."""
import ctypes as C
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import variogram as ov
from oracle.fields import gaussian_field, random_points

pytestmark = pytest.mark.gpu

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "scikit-gstat_b200", "skgstat_b200", "libskg_b200.so")


def _bind():
    lib = C.CDLL(LIB)
    P, I64, I32 = C.c_void_p, C.c_int64, C.c_int
    lib.skg_pair_hist.restype = I32
    lib.skg_pair_hist.argtypes = [P, P, I32, I32, I64, I32, P, P, I32, I32, P, P, P, P, P, I64, I64, I64, I64, P]
    lib.skg_pair_scratch_bytes.restype = I64
    lib.skg_pair_scratch_bytes.argtypes = [I64, I32]
    lib.skg_pair_num_tiles.restype = I64
    lib.skg_pair_num_tiles.argtypes = [I64]
    lib.skg_morton_keys.restype = I32
    lib.skg_morton_keys.argtypes = [P, I64, I32, P, P, P, P]
    lib.skg_last_error.restype = C.c_char_p
    return lib


def _experimental_matheron(lib, coords, values, edges, sort):
    import torch
    from skgstat_b200._exact import sq_threshold_bits
    n, dim = coords.shape
    xy = torch.from_numpy(np.ascontiguousarray(coords.T, dtype=np.float64)).cuda()
    z = torch.from_numpy(np.asarray(values, dtype=np.float64)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    if sort:
        lo = np.ascontiguousarray(np.nanmin(coords, axis=0))
        ext = np.ascontiguousarray(np.nanmax(coords, axis=0) - lo)
        keys = torch.empty(n, dtype=torch.int64, device="cuda")
        rc = lib.skg_morton_keys(xy.data_ptr(), n, dim, lo.ctypes.data, ext.ctypes.data,
                                 keys.data_ptr(), st)
        assert rc == 0, lib.skg_last_error()
        perm = torch.argsort(keys)
        xy = xy[:, perm].contiguous()
        z = z[perm].contiguous()
    thr = sq_threshold_bits(edges)
    count = torch.zeros(len(edges), dtype=torch.int64, device="cuda")
    ssq = torch.zeros(len(edges), dtype=torch.float64, device="cuda")
    scratch = torch.empty(lib.skg_pair_scratch_bytes(n, len(edges)), dtype=torch.uint8, device="cuda")
    rc = lib.skg_pair_hist(xy.data_ptr(), z.data_ptr(), 1, 0, n, dim, None, thr.ctypes.data, len(edges),
                           1, None, count.data_ptr(), ssq.data_ptr(), None, scratch.data_ptr(),
                           scratch.numel(), 0, lib.skg_pair_num_tiles(n), 1, st)
    assert rc == 0, lib.skg_last_error()
    c = count.cpu().numpy()
    with np.errstate(invalid="ignore", divide="ignore"):
        gamma = np.where(c == 0, np.nan, ssq.cpu().numpy() / (2 * c.astype(float)))
    return gamma, c


@pytest.mark.parametrize("sort", [False, True])
@pytest.mark.parametrize("n,dim", [(1500, 2), (1100, 3)])
def test_documented_stub_matches_oracle(n, dim, sort):
    lib = _bind()
    coords = random_points(n, 100.0, dim, seed=7)
    values = gaussian_field(coords, 15.0, seed=7)
    ref = ov.dense_variogram(coords, values, "matheron", "even", 12, "median")
    gamma, count = _experimental_matheron(lib, coords, values, ref["bins"], sort)
    assert_array_equal(count, ref["bin_count"])
    assert_allclose(gamma, ref["experimental"], rtol=1e-12, equal_nan=True)


def test_stub_rejects_bad_arguments_without_aborting():
    """Return-code convention of the header: negative SKG_E_*, text in skg_last_error()."""
    import torch
    lib = _bind()
    xy = torch.zeros(2, 8, dtype=torch.float64, device="cuda")
    z = torch.zeros(8, dtype=torch.float64, device="cuda")
    thr = np.array([5, 3], dtype=np.uint64)         # decreasing thresholds
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    scratch = torch.empty(lib.skg_pair_scratch_bytes(8, 2), dtype=torch.uint8, device="cuda")
    rc = lib.skg_pair_hist(xy.data_ptr(), z.data_ptr(), 1, 0, 8, 2, None, thr.ctypes.data, 2, 0, None,
                           cnt.data_ptr(), None, None, scratch.data_ptr(), scratch.numel(), 0, 1, 1, None)
    assert rc < 0
    assert b"non-decreasing" in lib.skg_last_error()
    rc = lib.skg_pair_hist(xy.data_ptr(), z.data_ptr(), 1, 0, 8, 5, None, thr.ctypes.data, 0, 0, None,
                           cnt.data_ptr(), None, None, scratch.data_ptr(), scratch.numel(), 0, 1, 1, None)
    assert rc < 0
