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


# ------------------------------------------------------------------------------------------------
# Sharded hash join: the exchange logic (three all-to-alls + order restoration) on CPU over gloo.
# The four data-path steps are CUDA kernels in the product (CudaJoinKernels); here a numpy stand-in
# with the same contract takes their place, and the concatenated per-rank results must equal the
# oracle's HashJoin over the whole tables, in order.
class NumpyJoinKernels(object):
    def scope(self):
        import contextlib
        return contextlib.nullcontext()

    def finish(self):
        pass

    @staticmethod
    def _part_of(keys, n_parts):
        h = np.zeros(keys[0][0].numel(), dtype=np.uint64)
        for t, _ in keys:
            h = h * np.uint64(1000003) + t.numpy().astype(np.int64).view(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        return ((h >> np.uint64(40)) % np.uint64(n_parts)).astype(np.int64)

    def partition(self, keys, n_parts, null_part):
        part = self._part_of(keys, n_parts)
        perm = np.argsort(part, kind="stable")
        return torch.from_numpy(perm), np.bincount(part, minlength=n_parts).tolist()

    def build_table(self, build_keys):
        """Stand-in for the table image: (key, row) pairs as one int64 tensor."""
        keys = build_keys[0][0].numpy().astype(np.int64)
        t = np.empty(2 * len(keys), dtype=np.int64)
        t[0::2], t[1::2] = keys, np.arange(len(keys))
        return torch.from_numpy(t), None

    def probe_parts(self, tables, row_offsets, probe_keys, join_type):
        index = []
        for t, off in zip(tables, row_offsets):
            a = t.numpy()
            d = {}
            for key, row in zip(a[0::2].tolist(), a[1::2].tolist()):
                d.setdefault(key, row + off)
            index.append(d)
        part = self._part_of(probe_keys, len(tables))
        li, ri = [], []
        for r, (key, p_) in enumerate(zip(probe_keys[0][0].numpy().tolist(), part.tolist())):
            m = index[p_].get(key)
            if m is not None:
                li.append(r)
                ri.append(m)
            elif join_type == 1:
                li.append(r)
                ri.append(-1)
        return torch.tensor(li, dtype=torch.int64), torch.tensor(ri, dtype=torch.int64)

    def gather(self, col, idx, want_valid=False):
        t, _ = col
        i = idx.numpy()
        src = t.numpy()
        out = np.zeros(len(i), dtype=src.dtype)
        ok = i >= 0
        out[ok] = src[i[ok]]
        if want_valid:
            return torch.from_numpy(out), torch.from_numpy((~ok).astype(np.uint8))
        return torch.from_numpy(out)

    def join(self, build_keys, probe_keys, join_type, uniqueness):
        index = {}
        bk = list(zip(*[t.numpy().tolist() for t, _ in build_keys])) if build_keys[0][0].numel() else []
        for r, key in enumerate(bk):
            index.setdefault(key, []).append(r)
        li, ri = [], []
        pk = list(zip(*[t.numpy().tolist() for t, _ in probe_keys])) if probe_keys[0][0].numel() else []
        for r, key in enumerate(pk):
            m = index.get(key)
            if m:
                for b in (m[:1] if uniqueness == 1 else m):
                    li.append(r)
                    ri.append(b)
            elif join_type == 1:
                li.append(r)
                ri.append(-1)
        return torch.tensor(li, dtype=torch.int64), torch.tensor(ri, dtype=torch.int64)

    def order_by(self, key):
        return torch.from_numpy(np.argsort(key.numpy(), kind="stable"))

    def sort_perm(self, keys, descending, key_nulls=None):
        ranks = []
        for i, ((t, dt), d) in enumerate(zip(keys, descending)):
            v = t.numpy()
            v = v.view(np.uint64) if dt == 3 else v.view(np.uint32) if dt == 8 else v      # unsigned columns: signed bit image
            isn = key_nulls[i].numpy().astype(bool) if key_nulls and key_nulls[i] is not None else np.zeros(len(v), dtype=bool)
            r = np.unique(np.where(isn, v[~isn][0] if (~isn).any() else 0, v), return_inverse=True)[1].astype(np.int64)
            r = -r if d else r
            if len(r):
                r[isn] = (r.max() + 1) if d else (r.min() - 1)      # NULLs last for DESC, first for ASC
            ranks.append(r)
        if not len(ranks[0]):
            return torch.zeros(0, dtype=torch.int64)
        return torch.from_numpy(np.lexsort(ranks[::-1]).astype(np.int64))   # lexsort is stable, last key = primary

    def scatter(self, col, idx, dst):
        dst.numpy()[idx.numpy()] = col[0].numpy()

    def zeros(self, n, dtype):
        from supersonic_b200.distributed import _torch_dtype
        return torch.zeros(int(n), dtype=_torch_dtype(dtype))

    def iota(self, n):
        return torch.arange(int(n), dtype=torch.int64)

    def compact(self, flag, cols):
        m = flag.numpy() != 0
        return [(torch.from_numpy(np.ascontiguousarray(t.numpy()[m])), dt) for t, dt in cols]


def _join_tables(uniq, scale=1):
    rng = np.random.default_rng(5)
    nb, npr = 3000 * scale, 20011 * scale
    pk = rng.permutation(nb).astype(np.int64) * 3
    if not uniq:
        pk[rng.integers(0, nb, nb // 5)] = pk[rng.integers(0, nb, nb // 5)]
    return {"pk": pk, "payload": rng.integers(0, 10**9, nb), "w": rng.random(nb),
            "fk": rng.integers(0, nb * 3, npr), "lv": rng.integers(0, 10**9, npr),
            "pk_null": (rng.random(nb) < 0.05).astype(np.uint8), "fk_null": (rng.random(npr) < 0.1).astype(np.uint8),
            "w_null": (rng.random(nb) < 0.2).astype(np.uint8)}


def _join_worker(rank, world, port, out, strategy):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from supersonic_b200.distributed import ShardedHashJoin
        I64, F64 = 2, 5
        res = {}
        for uniq in (1, 0):
            t = _join_tables(uniq)
            bb, be = shard_rows(len(t["pk"]), rank, world, align=1)
            pb, pe = shard_rows(len(t["fk"]), rank, world, align=1)
            col = lambda name, b, e, dt: (torch.from_numpy(np.ascontiguousarray(t[name][b:e])), dt)   # noqa: E731
            for jt in (0, 1):
                j = ShardedHashJoin(NumpyJoinKernels(), strategy=strategy)
                rows, lcols, rcols, rnull = j.run([col("fk", pb, pe, I64)], [col("fk", pb, pe, I64), col("lv", pb, pe, I64)],
                                                  [col("pk", bb, be, I64)], [col("payload", bb, be, I64), col("w", bb, be, F64)],
                                                  join_type=jt, uniqueness=uniq)
                res[(uniq, jt)] = ([c.numpy() for c, _ in lcols], [c.numpy() for c, _ in rcols],
                                   None if rnull is None else rnull.numpy(), rows.numpy() + pb)
        out.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("strategy", ["all_to_all", "broadcast", "replicate"])
def test_sharded_hash_join_world2_matches_oracle(ref, strategy):
    from supersonic_b200 import ssplan as sp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_join_worker, args=(r, 2, port, out, strategy)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for uniq in (1, 0):
        t = _join_tables(uniq)
        build = [sp.Column("pk", sp.INT64, t["pk"]), sp.Column("payload", sp.INT64, t["payload"]),
                 sp.Column("w", sp.DOUBLE, t["w"])]
        probe = [sp.Column("fk", sp.INT64, t["fk"]), sp.Column("lv", sp.INT64, t["lv"])]
        for jt in (0, 1):
            plan = ("(hash_join %s (named fk) (named pk) (multi (0 (all)) (1 (named payload w))) %s (scan 0) (scan 1))"
                    % (["INNER", "LEFT_OUTER"][jt], ["NOT_UNIQUE", "UNIQUE"][uniq]))
            want = ref.run(plan, [probe, build])
            assert want.code == 0
            parts = [got[r][(uniq, jt)] for r in range(2)]
            fk = np.concatenate([p[0][0] for p in parts])
            lv = np.concatenate([p[0][1] for p in parts])
            pay = np.concatenate([p[1][0] for p in parts])
            w = np.concatenate([p[1][1] for p in parts])
            assert len(fk) == want.rows
            assert np.array_equal(fk, want.columns[0]) and np.array_equal(lv, want.columns[1])
            if jt == 1:
                isn = np.concatenate([p[2] for p in parts]).astype(bool)
                assert np.array_equal(isn, want.nulls[2]) and np.array_equal(isn, want.nulls[3])
                assert np.array_equal(pay[~isn], want.columns[2][~isn]) and np.array_equal(w[~isn], want.columns[3][~isn])
            else:
                assert np.array_equal(pay, want.columns[2]) and np.array_equal(w, want.columns[3])
            # global lhs row ids are ascending: the concatenation is in lhs order
            gl = np.concatenate([p[3] for p in parts])
            assert np.all(np.diff(gl) >= 0)


def _null_join_worker(rank, world, port, out, strategy, use_cuda=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if use_cuda:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from supersonic_b200.distributed import ShardedHashJoin
        if use_cuda:
            from supersonic_b200 import capi
            from supersonic_b200.distributed import CudaJoinKernels
            kern = CudaJoinKernels(capi.Context(rank))
        else:
            kern = NumpyJoinKernels()
        I64, F64, BOOL = 2, 5, 6
        res = {}
        for uniq in (1, 0):
            t = _join_tables(uniq)
            bb, be = shard_rows(len(t["pk"]), rank, world, align=1)
            pb, pe = shard_rows(len(t["fk"]), rank, world, align=1)

            def col(name, b, e, dt):
                x = torch.from_numpy(np.ascontiguousarray(t[name][b:e]))
                return (x.cuda() if use_cuda else x, dt)
            for jt in (0, 1):
                j = ShardedHashJoin(kern, strategy=strategy)
                rows, lcols, rcols, rnull = j.run(
                    [col("fk", pb, pe, I64)], [col("lv", pb, pe, I64)],
                    [col("pk", bb, be, I64)], [col("payload", bb, be, I64), col("w", bb, be, F64), col("w_null", bb, be, BOOL)],
                    join_type=jt, uniqueness=uniq, lhs_key_nulls=col("fk_null", pb, pe, BOOL)[0],
                    rhs_key_nulls=col("pk_null", bb, be, BOOL)[0])
                res[(uniq, jt)] = ([c.cpu().numpy() for c, _ in lcols], [c.cpu().numpy() for c, _ in rcols],
                                   None if rnull is None else rnull.cpu().numpy(), rows.cpu().numpy() + pb)
        out.put((rank, res))
    finally:
        dist.destroy_process_group()


def check_null_join_against_oracle(ref, got):
    from supersonic_b200 import ssplan as sp
    for uniq in (1, 0):
        t = _join_tables(uniq)
        build = [sp.Column("pk", sp.INT64, t["pk"], is_null=t["pk_null"].astype(bool)), sp.Column("payload", sp.INT64, t["payload"]),
                 sp.Column("w", sp.DOUBLE, t["w"], is_null=t["w_null"].astype(bool))]
        probe = [sp.Column("fk", sp.INT64, t["fk"], is_null=t["fk_null"].astype(bool)), sp.Column("lv", sp.INT64, t["lv"])]
        for jt in (0, 1):
            plan = ("(hash_join %s (named fk) (named pk) (multi (0 (named lv)) (1 (named payload w))) %s (scan 0) (scan 1))"
                    % (["INNER", "LEFT_OUTER"][jt], ["NOT_UNIQUE", "UNIQUE"][uniq]))
            want = ref.run(plan, [probe, build])
            assert want.code == 0, want.error
            parts = [got[r][(uniq, jt)] for r in range(len(got))]
            lv = np.concatenate([p[0][0] for p in parts])
            pay = np.concatenate([p[1][0] for p in parts])
            w = np.concatenate([p[1][1] for p in parts])
            wn = np.concatenate([p[1][2] for p in parts]).astype(bool)
            assert len(lv) == want.rows and np.array_equal(lv, want.columns[0])
            miss = np.concatenate([p[2] for p in parts]).astype(bool) if jt == 1 else np.zeros(len(lv), dtype=bool)
            if jt == 1:
                assert np.array_equal(miss, want.nulls[1])
            assert np.array_equal(pay[~miss], want.columns[1][~miss])
            w_null = miss | wn
            assert np.array_equal(w_null, want.nulls[2] if want.nulls[2] is not None else np.zeros(len(lv), dtype=bool))
            assert np.array_equal(w[~w_null], want.columns[2][~w_null])


@pytest.mark.parametrize("strategy", ["all_to_all", "broadcast", "replicate"])
def test_sharded_hash_join_null_keys_and_payload_world2(ref, strategy):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_null_join_worker, args=(r, 2, port, out, strategy)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    check_null_join_against_oracle(ref, got)


def _sort_table(n=30011, flavour=""):
    rng = np.random.default_rng(11)
    if flavour == "one_value":     # every row has the same leading key: one rank receives everything
        return {"k": np.full(n, 7, dtype=np.int64), "x": np.round(rng.standard_normal(n), 1),
                "id": np.arange(n, dtype=np.int64), "v": np.full(n, 3, dtype=np.int64), "u": np.full(n, -5, dtype=np.int64)}
    if flavour == "empty":
        z = np.zeros(0, dtype=np.int64)
        return {"k": z, "x": np.zeros(0), "id": z, "v": z, "u": z}
    return {"k": rng.integers(-50, 50, n), "x": np.round(rng.standard_normal(n), 1), "id": np.arange(n, dtype=np.int64),
            "v": rng.integers(0, 10**9, n),
            # UINT64 values on both sides of 2^63, carried as their int64 bit image
            "u": (rng.integers(0, 200, n).astype(np.uint64) * np.uint64(1 << 57)).view(np.int64)}


SORT_CASES = [([("k", 2)], [False]), ([("k", 2)], [True]), ([("x", 5), ("k", 2)], [True, False]),
              ([("k", 2), ("x", 5)], [False, True]), ([("v", 2)], [False]), ([("u", 3)], [False]), ([("u", 3), ("k", 2)], [True, True])]


def _sort_worker(rank, world, port, out, skew, use_cuda=False, flavour=""):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if use_cuda:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from supersonic_b200.distributed import ShardedSort
        if use_cuda:
            from supersonic_b200 import capi
            from supersonic_b200.distributed import CudaJoinKernels
            kern = CudaJoinKernels(capi.Context(rank))
        else:
            kern = NumpyJoinKernels()
        place = (lambda x: x.cuda()) if use_cuda else (lambda x: x)
        t = _sort_table(flavour=flavour)
        n = len(t["k"])
        if skew:   # rank 0 holds nothing, the last rank holds most rows
            cuts = [0, 0] + [n // 7 * i for i in range(1, world - 1)] + [n]
            b, e = cuts[rank], cuts[rank + 1]
        else:
            b, e = shard_rows(n, rank, world, align=1)
        col = lambda name, dt: (place(torch.from_numpy(np.ascontiguousarray(t[name][b:e]))), dt)   # noqa: E731
        res = []
        for keys, desc in SORT_CASES:
            ks, cs = ShardedSort(kern).run([col(nm, dt) for nm, dt in keys], desc, [col("id", 2), col("v", 2)])
            res.append(([c.cpu().numpy() for c, _ in ks], [c.cpu().numpy() for c, _ in cs]))
        out.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,skew", [(2, False), (3, True)])
def test_sharded_sort_matches_oracle(ref, world, skew):
    """The concatenation of the ranks' ranges is the reference's Sort of the whole table; equal keys keep input
    order, which the plan pins by ordering on the row id last."""
    check_sharded_sort_against_oracle(ref, world, skew, False)


@pytest.mark.parametrize("flavour", ["one_value", "empty"])
def test_sharded_sort_degenerate_inputs(ref, flavour):
    check_sharded_sort_against_oracle(ref, 2, False, False, flavour)


def check_sharded_sort_against_oracle(ref, world, skew, use_cuda, flavour=""):
    from supersonic_b200 import ssplan as sp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sort_worker, args=(r, world, port, out, skew, use_cuda, flavour)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    t = _sort_table(flavour=flavour)
    table = [sp.Column("k", sp.INT64, t["k"]), sp.Column("x", sp.DOUBLE, t["x"]), sp.Column("id", sp.INT64, t["id"]),
             sp.Column("v", sp.INT64, t["v"]), sp.Column("u", sp.UINT64, t["u"].view(np.uint64))]
    for ci, (keys, desc) in enumerate(SORT_CASES):
        order = " ".join("(%s %s)" % (nm, "DESC" if d else "ASC") for (nm, _), d in zip(keys, desc)) + " (id ASC)"
        want = ref.run("(sort (order %s) (all) (scan 0))" % order, [table])
        assert want.code == 0, want.error
        ids = np.concatenate([got[r][ci][1][0] for r in range(world)])
        vs = np.concatenate([got[r][ci][1][1] for r in range(world)])
        assert np.array_equal(ids, want.column("id")) and np.array_equal(vs, want.column("v"))
        for j in range(len(keys)):
            kj = np.concatenate([got[r][ci][0][j] for r in range(world)])
            assert np.array_equal(kj, want.column(keys[j][0]).view(kj.dtype))
        if not skew and not flavour and keys[0][0] == "v":   # distinct keys: the ranges are balanced within a sampling error
            sizes = [len(got[r][ci][1][0]) for r in range(world)]
            assert max(sizes) < 1.2 * len(ids) / world


def _null_sort_table(n=20011):
    rng = np.random.default_rng(23)
    return {"k": rng.integers(-20, 20, n), "k_null": (rng.random(n) < 0.15).astype(np.uint8),
            "j": rng.integers(0, 5, n), "j_null": (rng.random(n) < 0.3).astype(np.uint8),
            "id": np.arange(n, dtype=np.int64), "w": rng.random(n), "w_null": (rng.random(n) < 0.2).astype(np.uint8)}


NULL_SORT_CASES = [[False, False], [True, False], [False, True], [True, True]]


def _null_sort_worker(rank, world, port, out, use_cuda=False, all_null=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if use_cuda:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from supersonic_b200.distributed import ShardedSort
        if use_cuda:
            from supersonic_b200 import capi
            from supersonic_b200.distributed import CudaJoinKernels
            kern = CudaJoinKernels(capi.Context(rank))
        else:
            kern = NumpyJoinKernels()
        place = (lambda x: x.cuda()) if use_cuda else (lambda x: x)
        t = _null_sort_table()
        if all_null:
            t["k_null"][:] = 1
        b, e = shard_rows(len(t["k"]), rank, world, align=1)
        ten = lambda name: place(torch.from_numpy(np.ascontiguousarray(t[name][b:e])))   # noqa: E731
        res = []
        for desc in NULL_SORT_CASES:
            ks, cs, kn, cn = ShardedSort(kern).run([(ten("k"), 2), (ten("j"), 2)], desc, [(ten("id"), 2), (ten("w"), 5)],
                                                   key_nulls=[ten("k_null"), ten("j_null")], col_nulls=[None, ten("w_null")])
            res.append(([c.cpu().numpy() for c, _ in ks], [c.cpu().numpy() for c, _ in cs],
                        [f.cpu().numpy() for f in kn], cn[1].cpu().numpy()))
        out.put((rank, res))
    finally:
        dist.destroy_process_group()


def check_null_sort_against_oracle(ref, world, use_cuda, all_null=False):
    from supersonic_b200 import ssplan as sp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_null_sort_worker, args=(r, world, port, out, use_cuda, all_null)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    t = _null_sort_table()
    if all_null:
        t["k_null"][:] = 1
    table = [sp.Column("k", sp.INT64, t["k"], t["k_null"].astype(bool)), sp.Column("j", sp.INT64, t["j"], t["j_null"].astype(bool)),
             sp.Column("id", sp.INT64, t["id"]), sp.Column("w", sp.DOUBLE, t["w"], t["w_null"].astype(bool))]
    for ci, desc in enumerate(NULL_SORT_CASES):
        want = ref.run("(sort (order (k %s) (j %s) (id ASC)) (all) (scan 0))" % tuple("DESC" if d else "ASC" for d in desc), [table])
        assert want.code == 0, want.error
        cat = lambda f: np.concatenate([f(got[r][ci]) for r in range(world)])   # noqa: E731
        assert np.array_equal(cat(lambda g: g[1][0]), want.column("id"))
        kn, jn, wn = cat(lambda g: g[2][0]).astype(bool), cat(lambda g: g[2][1]).astype(bool), cat(lambda g: g[3]).astype(bool)
        assert np.array_equal(kn, want.null("k")) and np.array_equal(jn, want.null("j")) and np.array_equal(wn, want.null("w"))
        assert np.array_equal(cat(lambda g: g[0][0])[~kn], want.column("k")[~kn])
        assert np.array_equal(cat(lambda g: g[0][1])[~jn], want.column("j")[~jn])
        assert np.array_equal(cat(lambda g: g[1][1])[~wn], want.column("w")[~wn])
        # NULL leading keys sit on the first rank for ASC and on the last for DESC
        holder = world - 1 if desc[0] else 0
        for r in range(world):
            assert r == holder or not got[r][ci][2][0].any()


@pytest.mark.parametrize("world,all_null", [(2, False), (3, False), (2, True)])
def test_sharded_sort_null_keys_match_oracle(ref, world, all_null):
    """Nullable key columns (NULLs first for ASC, last for DESC, per key) and a nullable payload column; with every
    leading key NULL no splitter exists and one rank collects the table."""
    check_null_sort_against_oracle(ref, world, False, all_null)
