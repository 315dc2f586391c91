"""Host-side logic of the multi-GPU path on CPU: row-range sharding and
the ragged all-gather of partial aggregate tables, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from supersonic_b200.distributed import allgather_ragged, shard_rows


def test_shard_rows_cover_and_align():
    for total in [0, 1, 1023, 1024, 1025, 10**6 + 7, 10**9]:
        for world in [1, 2, 3, 8]:
            prev = 0
            for r in range(world):
                b, e = shard_rows(total, r, world)
                assert b == prev and e >= b
                if r + 1 < world:
                    assert e % 1024 == 0 or e == total
                prev = e
            assert prev == total


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank aggregates its row range on CPU here (the GPU kernel is covered by the -m gpu
        # tests); what is under test is the exchange: ragged partial tables must arrive complete and
        # in rank order, and merging them must give the whole-table answer.
        b, e = shard_rows(total, rank, world)
        rng = np.random.default_rng(1)
        keys = rng.integers(0, 50 + 37 * 0, total)
        vals = rng.integers(0, 1000, total).astype(np.float64)
        k, v = keys[b:e], vals[b:e]
        uk = np.unique(k)
        sums = np.array([v[k == x].sum() for x in uk])
        cnts = np.array([(k == x).sum() for x in uk], dtype=np.int64)
        if rank == 1:
            uk, sums, cnts = uk[:-3], sums[:-3], cnts[:-3]   # make the partial tables ragged
        gk = allgather_ragged(torch.from_numpy(uk))
        gs = allgather_ragged(torch.from_numpy(sums))
        gc = allgather_ragged(torch.from_numpy(cnts))
        assert [len(x) for x in gk] == [len(x) for x in gs] == [len(x) for x in gc]
        merged = {}
        for kk, ss, cc in zip(gk, gs, gc):
            for a, s_, c_ in zip(kk.tolist(), ss.tolist(), cc.tolist()):
                m = merged.setdefault(a, [0.0, 0])
                m[0] += s_
                m[1] += c_
        if rank == 0:
            out.put(sorted((a, m[0], m[1]) for a, m in merged.items()))
    finally:
        dist.destroy_process_group()


def test_ragged_allgather_and_merge_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    total = 20000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected: whole table minus the three groups rank 1 dropped from its partial
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 50, total)
    vals = rng.integers(0, 1000, total).astype(np.float64)
    b1, e1 = shard_rows(total, 1, 2)
    dropped = np.unique(keys[b1:e1])[-3:]
    want = []
    for x in np.unique(keys):
        m0 = keys[:b1] == x
        s, c = vals[:b1][m0].sum(), int(m0.sum())
        if x not in dropped:
            m1 = keys[b1:e1] == x
            s, c = s + vals[b1:e1][m1].sum(), c + int(m1.sum())
        want.append((int(x), float(s), c))
    assert got == want
