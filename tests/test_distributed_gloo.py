"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (tile sharding, the single
all-reduce of the per-lag partials, histogram all-reduce and candidate all-gather of the
exact select) gives the single-process answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from skgstat_b200 import distributed as D
from skgstat_b200._exact import to_bits
from skgstat_b200.select import exact_select


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert D.world() == (rank, world)
        # --- fused pass: each rank's partial [count | sum_sq | sum_sqrt] meets in ONE all-reduce
        rng = np.random.default_rng(0)
        tiles = 37
        per_tile = rng.random((tiles, 3 * 5))
        per_tile[:, :5] = np.floor(per_tile[:, :5] * 1000)
        b, e = D.shard_range(tiles, rank, world)
        part = torch.from_numpy(per_tile[b:e].sum(axis=0))
        D.allreduce_sum_(part)
        np.testing.assert_allclose(part.numpy(), per_tile.sum(axis=0), rtol=1e-13)
        assert np.array_equal(part.numpy()[:5], per_tile[:, :5].sum(axis=0))  # counts exact
        # --- max
        m = torch.tensor([int(rank + 5)], dtype=torch.int64)
        D.allreduce_max_(m)
        assert int(m) == world + 4
        # --- varlen gather keeps rank order
        g = D.allgather_varlen(torch.arange(rank + 2, dtype=torch.float64) + 10 * rank)
        assert g.tolist() == [0.0, 1.0, 10.0, 11.0, 12.0]
        # --- exact select over keys sharded by "tile range"
        keys = np.floor(np.random.default_rng(1).random(30000) * 5000.0) / 7.0
        kb, ke = D.shard_range(keys.size, rank, world)
        mine = keys[kb:ke]
        nb = 128

        def buckets(lo, hi, sc):
            kbits = mine.view(np.uint64)
            msk = (kbits >= lo[0]) & (kbits < hi[0])
            lo_val = np.array([lo[0]], dtype=np.uint64).view(np.float64)[0]
            t = (mine[msk] - lo_val) * sc[0]
            return mine[msk], np.minimum(nb - 1, t.astype(np.int64))

        def run_hist(lo, hi, sc):
            _, bk = buckets(lo, hi, sc)
            h = torch.from_numpy(np.bincount(bk, minlength=nb).astype(np.int64))
            return D.allreduce_sum_(h).numpy()

        def run_gather(lo, hi, sc, bitmap, cap):
            k, bk = buckets(lo, hi, sc)
            hit = ((bitmap[bk >> 5] >> (bk & 31).astype(np.uint32)) & 1).astype(bool)
            kk = D.allgather_varlen(torch.from_numpy(k[hit]))
            ss = D.allgather_varlen(torch.from_numpy(bk[hit]))
            return kk.numpy(), ss.numpy()

        ranks = [0, 14999, 15000, 29999]
        vals, totals = exact_select(1, [0], [to_bits(float(keys.max())) + 1], nb,
                                    lambda g_, t_: ranks, run_hist, run_gather, bucket_limit=200)
        srt = np.sort(keys)
        assert totals[0] == keys.size
        for r in ranks:
            assert vals[0][r] == srt[r]
        out.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        out.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 3160, 12345):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_collectives_and_select():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res
