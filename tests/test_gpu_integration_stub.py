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


@pytest.mark.parametrize("m,cap,n_slots,decimals", [
    (1, 8, 1, 1), (257, 300, 1, 1), (5000, 5000, 3, 1), (700, 512, 2, 1), (0, 16, 1, 1),
    (100000, 100000, 3, 9),      # wide path: narrowing histogram, then ranking inside a bucket
    (300000, 250000, 1, 9),      # wide path, candidates beyond the capacity ignored
    (60000, 60000, 2, None),     # wide path, all keys equal: reports -1, the caller sorts
])
def test_pick_order_stats_against_numpy_partition(m, cap, n_slots, decimals):
    """skg_pick_order_stats: ranks by counting, ties by position, only the first min(count, capacity)
    candidates are looked at; unwanted slots (and the -1 filler) are ignored."""
    import torch
    lib = C.CDLL(LIB)
    P, I64, I32 = C.c_void_p, C.c_int64, C.c_int
    lib.skg_pick_order_stats.restype = I32
    lib.skg_pick_order_stats.argtypes = [P, P, P, I64, P, P, P, P, P, I32, P, P, I64, P]
    lib.skg_pick_scratch_bytes.restype = I64
    rng = np.random.default_rng(m + cap)
    keys = np.round(10.0 + rng.random(max(cap, 1)) * 50.0, decimals) if decimals is not None \
        else np.full(max(cap, 1), 33.0)
    slot_ids = np.array([7, 4096, 123456], dtype=np.uint32)[:n_slots]
    slots = rng.choice(np.concatenate((slot_ids, [99])), size=max(cap, 1)).astype(np.uint32)
    seen = min(m, cap)
    slots[seen:] = np.uint32(0xFFFFFFFF)
    want_slot, want_off, want_rank, expect_cnt, expect_val = [], [0], [], [], []
    for sid in slot_ids:
        mine = np.sort(keys[:seen][slots[:seen] == sid])
        ranks = sorted({0, mine.size // 2, max(mine.size - 1, 0), (mine.size - 1) // 2}) if mine.size else [0]
        want_slot.append(sid)
        want_rank += ranks
        want_off.append(len(want_rank))
        expect_cnt.append(mine.size)
        expect_val += [mine[r] if r < mine.size else 0.0 for r in ranks]
    ws = np.array(want_slot, dtype=np.uint32)
    wo = np.array(want_off, dtype=np.int32)
    wr = np.array(want_rank, dtype=np.int64)
    lo = np.full(len(ws), np.float64(10.0).view(np.uint64), dtype=np.uint64)
    hi = np.full(len(ws), np.float64(60.5).view(np.uint64), dtype=np.uint64)
    k_d = torch.from_numpy(keys).cuda()
    s_d = torch.from_numpy(slots.view(np.int32)).cuda()
    c_d = torch.tensor([m], dtype=torch.int64, device="cuda")
    out = torch.full((len(ws) + len(wr),), -7.0, dtype=torch.float64, device="cuda")
    scr = torch.empty(lib.skg_pick_scratch_bytes(), dtype=torch.uint8, device="cuda")
    rc = lib.skg_pick_order_stats(k_d.data_ptr(), s_d.data_ptr(), c_d.data_ptr(), cap, ws.ctypes.data,
                                  wo.ctypes.data, wr.ctypes.data, lo.ctypes.data, hi.ctypes.data, len(ws),
                                  out.data_ptr(), scr.data_ptr(), scr.numel(),
                                  torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    got = out.cpu().numpy()
    if decimals is None:
        assert np.all(got[:len(ws)] == -1.0)     # one bucket holds every key of the slot: list overflow
        return
    assert_array_equal(got[:len(ws)], np.array(expect_cnt, dtype=np.float64))
    assert_array_equal(got[len(ws):], np.array(expect_val))
