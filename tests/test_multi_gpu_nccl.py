"""The sharded hash join and sort on real GPUs: two ranks over NCCL, CudaJoinKernels (every data-path
step a kernel of libssb200.so), checked in order against the oracle's HashJoin over the whole
tables. Needs two B200s (`gpurun --gpus 2`); skipped on a single-GPU box."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_multi_gpu import _free_port, _join_tables
from supersonic_b200.distributed import shard_rows

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out, n_scale, strategy):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from supersonic_b200 import capi
        from supersonic_b200.distributed import CudaJoinKernels, ShardedHashJoin
        ctx = capi.Context(rank)
        kern = CudaJoinKernels(ctx)
        I64, F64 = capi.INT64, capi.DOUBLE
        res = {}
        for uniq in (1, 0):
            t = _join_tables(uniq, n_scale)
            bb, be = shard_rows(len(t["pk"]), rank, world, align=1)
            pb, pe = shard_rows(len(t["fk"]), rank, world, align=1)
            col = lambda name, b, e, dt: (torch.from_numpy(np.ascontiguousarray(t[name][b:e])).cuda(), dt)   # noqa: E731
            for jt in (0, 1):
                j = ShardedHashJoin(kern, strategy=strategy)
                rows, lcols, rcols, rnull = j.run([col("fk", pb, pe, I64)], [col("fk", pb, pe, I64), col("lv", pb, pe, I64)],
                                                  [col("pk", bb, be, I64)], [col("payload", bb, be, I64), col("w", bb, be, F64)],
                                                  join_type=jt, uniqueness=uniq)
                res[(uniq, jt)] = ([c.cpu().numpy() for c, _ in lcols], [c.cpu().numpy() for c, _ in rcols],
                                   None if rnull is None else rnull.cpu().numpy(), rows.cpu().numpy() + pb)
        launches = ctx.launches()
        out.put((rank, res, launches))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("strategy", ["all_to_all", "broadcast", "replicate"])
@pytest.mark.parametrize("n_scale", [1, 60])
def test_sharded_hash_join_two_gpus_matches_oracle(ref, n_scale, strategy):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from supersonic_b200 import ssplan as sp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, n_scale, strategy)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, res, launches = out.get(timeout=600)
        assert launches > 0
        got[rank] = res
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for uniq in (1, 0):
        t = _join_tables(uniq, n_scale)
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
                assert np.array_equal(isn, want.nulls[2])
                assert np.array_equal(pay[~isn], want.columns[2][~isn]) and np.array_equal(w[~isn], want.columns[3][~isn])
            else:
                assert np.array_equal(pay, want.columns[2]) and np.array_equal(w, want.columns[3])


@pytest.mark.parametrize("strategy", ["all_to_all", "broadcast", "replicate"])
def test_sharded_hash_join_null_keys_and_payload_two_gpus(ref, strategy):
    """NULL keys on both sides and a nullable payload column through the real kernels over NCCL."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from test_multi_gpu import _null_join_worker, check_null_join_against_oracle
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_null_join_worker, args=(r, 2, port, out, strategy, True)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    check_null_join_against_oracle(ref, got)


@pytest.mark.parametrize("skew", [False, True])
def test_sharded_sort_two_gpus_matches_oracle(ref, skew):
    """Sample sort over NCCL with the library's sort and gather kernels, against the oracle's Sort of the whole table."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from test_multi_gpu import check_sharded_sort_against_oracle
    check_sharded_sort_against_oracle(ref, 2, skew, True)


def test_sharded_sort_null_keys_two_gpus_matches_oracle(ref):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from test_multi_gpu import check_null_sort_against_oracle
    check_null_sort_against_oracle(ref, 2, True)


# ------------------------------------------------------------------------------------------------
# ssb_comm_* / ssb_shard_group_merge: the C ABI's own exchange (NCCL inside libssb200.so)
def _group_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import ctypes as C
        from supersonic_b200 import capi
        from supersonic_b200.distributed import make_comm, merge_group_partials
        ctx = capi.Context(rank)
        comm = make_comm(ctx)
        t = _group_tables()
        n = len(t["k"])
        b, e = shard_rows(n, rank, world, align=32)
        res = {}
        for name, keys, aggs in _GROUP_PLANS:
            def up(col):
                arr = np.ascontiguousarray(t[col][b:e])
                ptr = ctx.malloc(arr.nbytes + 256)
                ctx.h2d(ptr, arr)
                nptr = None
                if col + "_null" in t:
                    nl = t[col + "_null"][b:e]
                    words = np.packbits(np.concatenate([nl.astype(np.uint8), np.zeros((-len(nl)) % 32 + 32, np.uint8)]), bitorder="little")
                    nptr = ctx.malloc(words.nbytes + 256)
                    ctx.h2d(nptr, words)
                return ptr, nptr
            DT = {"k": capi.INT64, "k2": capi.INT32, "v": capi.DOUBLE, "w": capi.INT64}
            NPT = {capi.INT64: np.int64, capi.INT32: np.int32, capi.DOUBLE: np.float64, capi.UINT64: np.uint64}
            specs = (capi.AggSpec * len(aggs))()
            vals, vidx = [], {}
            for i, (fn, col, out_t) in enumerate(aggs):
                specs[i].fn, specs[i].out_type = fn, out_t
                if col is None:
                    specs[i].input, specs[i].in_type = -1, capi.INT64
                else:
                    if col not in vidx:
                        vidx[col] = len(vals)
                        vals.append(col)
                    specs[i].input, specs[i].in_type, specs[i].in_nullable = vidx[col], DT[col], 1 if col + "_null" in t else 0
            kt = (C.c_int32 * max(1, len(keys)))(*[DT[c] for c in keys])
            kn = (C.c_int32 * max(1, len(keys)))(*[1 if c + "_null" in t else 0 for c in keys])
            g = C.c_void_p()
            ctx.check(ctx.lib.ssb_group_create(ctx.h, len(keys), kt, kn, len(aggs), specs, 0, C.byref(g)))
            kc = (capi.Column * max(1, len(keys)))()
            for i, c in enumerate(keys):
                p_, n_ = up(c)
                kc[i].data, kc[i].nulls, kc[i].dtype = p_, n_, DT[c]
            vc = (capi.Column * max(1, len(vals)))()
            for i, c in enumerate(vals):
                p_, n_ = up(c)
                vc[i].data, vc[i].nulls, vc[i].dtype = p_, n_, DT[c]
            ctx.check(ctx.lib.ssb_group_update(g, kc, vc, e - b))
            ng, ko, ao = merge_group_partials(ctx, g, [DT[c] for c in keys], [a[2] for a in aggs], comm=comm)

            def down(col, dt):
                a = np.empty(ng, dtype=NPT[dt])
                isn = np.zeros(ng, dtype=bool)
                if ng:
                    ctx.d2h(a, col.data)
                    if col.nulls:
                        w = np.empty((ng + 31) // 32, dtype=np.uint32)
                        ctx.d2h(w, col.nulls)
                        isn = np.unpackbits(w.view(np.uint8), bitorder="little")[:ng].astype(bool)
                return a, isn
            res[name] = ([down(ko[i], DT[c]) for i, c in enumerate(keys)], [down(ao[i], a[2]) for i, a in enumerate(aggs)])
            ctx.lib.ssb_group_destroy(g)
        comm.close()
        out.put((rank, res))
    finally:
        dist.destroy_process_group()


def _group_tables():
    rng = np.random.default_rng(99)
    n = 400_000
    k = rng.integers(0, 5000, n)
    return {"k": k, "k_null": rng.random(n) < 0.01,
            "k2": rng.integers(0, 3, n).astype(np.int32),
            "v": rng.integers(0, 1 << 20, n) / 1024.0, "v_null": (k % 7 == 0) | (rng.random(n) < 0.05),
            "w": rng.integers(-10**6, 10**6, n)}


# (name, key columns, [(fn, value column or None, out type)])
_GROUP_PLANS = [
    ("one_key", ["k"], [(0, "v", 5), (3, None, 3), (1, "w", 2), (2, "v", 5), (3, "v", 3)]),
    ("two_keys", ["k", "k2"], [(0, "w", 2), (3, None, 3)]),
    ("scalar", [], [(0, "w", 2), (2, "v", 5), (3, None, 3)]),
]


def test_shard_group_merge_two_gpus_matches_oracle(ref):
    """ssb_shard_group_merge (reduce-scatter by key hash + all-gather inside libssb200.so) over two ranks against the
    oracle's GroupAggregate over the whole table: NULL keys form a group, all-NULL inputs give NULL (the `v` of every
    seventh key is always NULL), every rank ends with the whole result."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from supersonic_b200 import ssplan as sp
    from cases import same_results  # noqa: F401
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_group_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in procs:
        rank, res = out.get(timeout=600)
        got[rank] = res
    for p in procs:
        p.join(timeout=60)
    t = _group_tables()
    # the value column's NULLs depend on the key so that some groups are all-NULL
    cols = [sp.Column("k", sp.INT64, t["k"], is_null=t["k_null"]), sp.Column("k2", sp.INT32, t["k2"]),
            sp.Column("v", sp.DOUBLE, t["v"], is_null=t["v_null"]), sp.Column("w", sp.INT64, t["w"])]
    FN = {0: "SUM", 1: "MIN", 2: "MAX", 3: "COUNT"}
    for name, keys, aggs in _GROUP_PLANS:
        spec = " ".join('(%s %s a%d)' % (FN[fn], '""' if col is None else col, i) for i, (fn, col, _) in enumerate(aggs))
        plan = ("(group (named %s) (aggs %s) (scan 0))" % (" ".join(keys), spec)) if keys else "(scalar_agg (aggs %s) (scan 0))" % spec
        want = ref.run(plan, [cols])
        assert want.code == 0, want.error
        for rank in (0, 1):
            kcols, acols = got[rank][name]
            n = len(acols[0][0])
            assert n == want.rows, (name, rank, n, want.rows)
            gv = [np.where(isn, 0, a) for a, isn in kcols + acols]
            gn = [isn for _, isn in kcols + acols]
            wv = [np.where(want.nulls[i] if want.nulls[i] is not None else False, 0, want.columns[i]) for i in range(len(want.columns))]
            wn = [want.nulls[i] if want.nulls[i] is not None else np.zeros(want.rows, bool) for i in range(len(want.columns))]
            nk = len(keys)
            og = np.lexsort([x for i in reversed(range(nk)) for x in (gv[i], gn[i])]) if nk else np.arange(n)
            ow = np.lexsort([x for i in reversed(range(nk)) for x in (wv[i], wn[i])]) if nk else np.arange(n)
            for i in range(len(gv)):
                assert np.array_equal(gn[i][og], wn[i][ow]), (name, rank, i, "null flags")
                assert np.array_equal(gv[i][og], wv[i][ow]), (name, rank, i, "values")


# ------------------------------------------------------------------------------------------------
# ssb_shard_join_*: the sharded HashJoin behind the C ABI (hash partition + one grouped exchange + per-part tables
# all-gathered inside libssb200.so; torch only carries the communicator id)
def _shard_join_worker(rank, world, port, out, n_scale):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import ctypes as C
        from supersonic_b200 import capi
        from supersonic_b200.distributed import make_comm
        ctx = capi.Context(rank)
        comm = make_comm(ctx)
        lib = ctx.lib
        t = _join_tables(1, n_scale)
        bb, be = shard_rows(len(t["pk"]), rank, world, align=1)
        pb, pe = shard_rows(len(t["fk"]), rank, world, align=1)

        def up(arr, nulls=None):
            arr = np.ascontiguousarray(arr)
            ptr = ctx.malloc(arr.nbytes + 256)
            if arr.nbytes:
                ctx.h2d(ptr, arr)
            nptr = None
            if nulls is not None:
                words = np.packbits(np.concatenate([nulls.astype(np.uint8), np.zeros((-len(nulls)) % 32 + 32, np.uint8)]), bitorder="little")
                nptr = ctx.malloc(words.nbytes + 256)
                ctx.h2d(nptr, words)
            return ptr, nptr

        def cols(specs):
            arr = (capi.Column * max(1, len(specs)))()
            for i, (ptr, nptr, dt) in enumerate(specs):
                arr[i].data, arr[i].nulls, arr[i].dtype = ptr, nptr, dt
            return arr
        res = {}
        for with_nulls in (False, True):
            pk = up(t["pk"][bb:be], t["pk_null"][bb:be] if with_nulls else None)
            pay, w = up(t["payload"][bb:be]), up(t["w"][bb:be])
            fk = up(t["fk"][pb:pe], t["fk_null"][pb:pe] if with_nulls else None)
            lvp = up(t["lv"][pb:pe])
            j = C.c_void_p()
            ctx.check(lib.ssb_shard_join_build(comm.h, cols([(pk[0], pk[1], capi.INT64)]), 2,
                                               cols([(pay[0], None, capi.INT64), (w[0], None, capi.DOUBLE)]), be - bb, C.byref(j)))
            for jt in (0, 1):
                n, pl, pr = C.c_int64(), C.c_void_p(), C.c_void_p()
                ctx.check(lib.ssb_shard_join_probe(j, cols([(fk[0], fk[1], capi.INT64)]), pe - pb, jt, C.byref(n), C.byref(pl), C.byref(pr)))
                li = np.empty(n.value, dtype=np.int64)
                ri = np.empty(n.value, dtype=np.int64)
                if n.value:
                    ctx.d2h(li, pl)
                    ctx.d2h(ri, pr)
                outs = []
                for i, npdt in enumerate((np.int64, np.float64)):
                    c = capi.Column()
                    total = C.c_int64()
                    ctx.check(lib.ssb_shard_join_payload(j, i, C.byref(c), C.byref(total)))
                    whole = np.empty(total.value, dtype=npdt)
                    if total.value:
                        ctx.d2h(whole, c.data)
                    outs.append(np.where(ri >= 0, whole[np.maximum(ri, 0)], 0))
                res[(with_nulls, jt)] = (li + pb, ri < 0, outs)
                # the materialising probe writes the same cells itself
                o_lv, o_pay, o_w = ctx.malloc((pe - pb) * 8 + 256), ctx.malloc((pe - pb) * 8 + 256), ctx.malloc((pe - pb) * 8 + 256)
                d_match = ctx.malloc((pe - pb) + 256)
                m = C.c_int64()
                which = (C.c_int32 * 2)(0, 1)
                ctx.check(lib.ssb_shard_join_probe_materialize(j, cols([(fk[0], fk[1], capi.INT64)]), pe - pb, jt, 1, cols([(lvp[0], None, capi.INT64)]),
                                                               2, which, cols([(o_lv, None, capi.INT64), (o_pay, None, capi.INT64), (o_w, None, capi.DOUBLE)]),
                                                               d_match, C.byref(m)))
                assert m.value == n.value
                for ptr, npdt, want in ((o_lv, np.int64, t["lv"][pb:pe][li]), (o_pay, np.int64, outs[0]), (o_w, np.float64, outs[1])):
                    a = np.empty(m.value, dtype=npdt)
                    if m.value:
                        ctx.d2h(a, ptr)
                    assert np.array_equal(a, want)
            lib.ssb_shard_join_destroy(j)
        launches = ctx.launches()
        comm.close()
        out.put((rank, res, launches))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_scale", [1, 60])
def test_c_abi_sharded_join_two_gpus_matches_oracle(ref, n_scale):
    """ssb_shard_join_build / _probe / _payload over two ranks against the oracle's HashJoin over the whole tables:
    pairs in lhs order, payload values bit-exact, NULL keys on both sides never match."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from supersonic_b200 import ssplan as sp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_join_worker, args=(r, 2, port, out, n_scale)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in procs:
        rank, res, launches = out.get(timeout=600)
        assert launches > 0
        got[rank] = res
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    t = _join_tables(1, n_scale)
    for with_nulls in (False, True):
        build = [sp.Column("pk", sp.INT64, t["pk"], is_null=t["pk_null"].astype(bool) if with_nulls else None),
                 sp.Column("payload", sp.INT64, t["payload"]), sp.Column("w", sp.DOUBLE, t["w"])]
        probe = [sp.Column("fk", sp.INT64, t["fk"], is_null=t["fk_null"].astype(bool) if with_nulls else None),
                 sp.Column("lv", sp.INT64, t["lv"])]
        for jt in (0, 1):
            plan = ("(hash_join %s (named fk) (named pk) (multi (0 (named lv)) (1 (named payload w))) UNIQUE (scan 0) (scan 1))"
                    % ["INNER", "LEFT_OUTER"][jt])
            want = ref.run(plan, [probe, build])
            assert want.code == 0, want.error
            li = np.concatenate([got[r][(with_nulls, jt)][0] for r in range(2)])
            miss = np.concatenate([got[r][(with_nulls, jt)][1] for r in range(2)])
            pay = np.concatenate([got[r][(with_nulls, jt)][2][0] for r in range(2)])
            w = np.concatenate([got[r][(with_nulls, jt)][2][1] for r in range(2)])
            assert len(li) == want.rows, (with_nulls, jt, len(li), want.rows)
            assert np.array_equal(t["lv"][li], want.columns[0])
            wn = want.nulls[1] if want.nulls[1] is not None else np.zeros(want.rows, bool)
            assert np.array_equal(miss, wn)
            assert np.array_equal(pay[~miss], want.columns[1][~miss]) and np.array_equal(w[~miss], want.columns[2][~miss])
