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


def _sorted_device_points(lib, coords, values):
    import torch
    n, dim = coords.shape
    xy = torch.from_numpy(np.ascontiguousarray(coords.T, dtype=np.float64)).cuda()
    z = torch.from_numpy(np.asarray(values, dtype=np.float64)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    lo = np.ascontiguousarray(np.nanmin(coords, axis=0))
    ext = np.ascontiguousarray(np.nanmax(coords, axis=0) - lo)
    keys = torch.empty(n, dtype=torch.int64, device="cuda")
    assert lib.skg_morton_keys(xy.data_ptr(), n, dim, lo.ctypes.data, ext.ctypes.data, keys.data_ptr(), st) == 0
    perm = torch.argsort(keys)
    return xy[:, perm].contiguous(), z[perm].contiguous(), st


@pytest.mark.parametrize("n", [3000, 24000])
def test_round2_entry_points_through_the_raw_abi(n):
    """Count-only sweep (bulk counts + register windows at the larger size, n = 24000) and the one-sweep Dowd
    pass (skg_pair_dz_windows) straight through ctypes against numpy on the materialised vectors."""
    import torch
    from skgstat_b200._exact import sq_threshold_bits, to_bits
    lib = _bind()
    P, I64, I32 = C.c_void_p, C.c_int64, C.c_int
    lib.skg_pair_dz_windows.restype = I32
    lib.skg_pair_dz_windows.argtypes = [P, P, I32, I32, I64, I32, P, P, I32, P, P, P, I64, P, P, P, P, I64, I64, I64,
                                        I64, P]
    coords = random_points(n, 1000.0, 2, seed=21)
    values = gaussian_field(coords, 120.0, seed=21)
    xy, z, st = _sorted_device_points(lib, coords, values)
    # reference on a row-blocked sweep (no O(N^2) array)
    edges = np.linspace(0, 480.0, 11)[1:]
    nl = len(edges)
    cnt_ref = np.zeros(nl, dtype=np.int64)
    dz_lo = np.full(nl, 0.8)
    dz_hi = np.full(nl, 1.6)
    below_ref = np.zeros(nl, dtype=np.int64)
    nb = 64
    hist_ref = np.zeros((nl, nb), dtype=np.int64)
    scale = nb / (np.nextafter(dz_hi, 0) - dz_lo)
    for i0 in range(0, n - 1, 512):
        i1 = min(n - 1, i0 + 512)
        d, dz, _, _ = ov._row_block(ov.coords_2d(coords), np.asarray(values, dtype=float), i0, i1)
        g = ov.lag_groups(d, edges)
        ok = g >= 0
        cnt_ref += np.bincount(g[ok], minlength=nl)
        below_ref += np.bincount(g[ok & (dz < dz_lo[0])], minlength=nl)
        inw = ok & (dz >= dz_lo[0]) & (dz < dz_hi[0])
        b = np.minimum(nb - 1, ((dz[inw] - dz_lo[0]) * scale[0]).astype(np.int64))
        np.add.at(hist_ref, (g[inw], b), 1)
    thr = sq_threshold_bits(edges)
    scratch = torch.empty(lib.skg_pair_scratch_bytes(n, nl), dtype=torch.uint8, device="cuda")
    ntiles = lib.skg_pair_num_tiles(n)
    # stats = 0: exact counts, no values needed
    count = torch.zeros(nl, dtype=torch.int64, device="cuda")
    rc = lib.skg_pair_hist(xy.data_ptr(), z.data_ptr(), 1, 0, n, 2, None, thr.ctypes.data, nl, 0, None,
                           count.data_ptr(), None, None, scratch.data_ptr(), scratch.numel(), 0, ntiles, 1, st)
    assert rc == 0, lib.skg_last_error()
    assert_array_equal(count.cpu().numpy(), cnt_ref)
    # one-sweep Dowd pass
    lo_bits = np.array([to_bits(x) for x in dz_lo], dtype=np.uint64)
    hi_bits = np.array([to_bits(x) for x in dz_hi], dtype=np.uint64)
    count.zero_()
    below = torch.zeros(nl, dtype=torch.float64, device="cuda")
    hist = torch.zeros(nl * nb, dtype=torch.int64, device="cuda")
    rc = lib.skg_pair_dz_windows(xy.data_ptr(), z.data_ptr(), 1, 0, n, 2, None, thr.ctypes.data, nl,
                                 lo_bits.ctypes.data, hi_bits.ctypes.data, scale.ctypes.data, nb, count.data_ptr(),
                                 below.data_ptr(), hist.data_ptr(), scratch.data_ptr(), scratch.numel(), 0, ntiles,
                                 1, st)
    assert rc == 0, lib.skg_last_error()
    assert_array_equal(count.cpu().numpy(), cnt_ref)
    assert_array_equal(below.cpu().numpy().astype(np.int64), below_ref)
    assert_array_equal(hist.cpu().numpy().reshape(nl, nb), hist_ref)
