"""Device-resident checks through the C ABI at sizes closer to BASELINE.json: the table is
generated in HBM by the counter-based generator, the host regenerates the same rows chunk by
chunk (ssb_generate_host is bit-identical) and verifies the kernels' outputs."""
import ctypes as C

import numpy as np
import pytest

from supersonic_b200 import capi

pytestmark = pytest.mark.gpu

GEN = {"a": (0, -(1 << 31), 1 << 32), "b": (0, -(1 << 31), 1 << 32), "c": (0, -(1 << 62), 1 << 63), "d": (0, 0, 1 << 20)}


@pytest.fixture(scope="module")
def ctx(built):
    c = capi.Context(0)
    yield c
    c.close()


def host_col(name, rows, first=0):
    kind, lo, span = GEN[name]
    out = np.empty(rows, dtype=np.int64)
    capi.load().ssb_generate_host(out.ctypes.data, rows, first, 42, "abcd".index(name), kind, lo, span)
    return out


def test_generator_device_equals_host(ctx):
    rows = 1_000_003
    for name in "abcd":
        kind, lo, span = GEN[name]
        p = ctx.malloc(rows * 8)
        ctx.generate(p, rows, 77, 42, "abcd".index(name), kind, lo, span)
        got = np.empty(rows, dtype=np.int64)
        ctx.d2h(got, p)
        want = np.empty(rows, dtype=np.int64)
        ctx.lib.ssb_generate_host(want.ctypes.data, rows, 77, 42, "abcd".index(name), kind, lo, span)
        assert np.array_equal(got, want)
        ctx.free(p)


def c2_program(ctx, k):
    n, I64, B = capi.node, capi.INT64, capi.BOOL
    nodes = [n(capi.OP_INPUT, I64, [0]), n(capi.OP_INPUT, I64, [1]), n(capi.OP_INPUT, I64, [2]), n(capi.OP_INPUT, I64, [3]),
             n(capi.OP_MUL, I64, [0, 1]), n(capi.OP_ADD, I64, [4, 2]), n(capi.OP_CONST, I64, [], i64=k), n(capi.OP_LT, B, [3, 6])]
    return capi.Program(ctx, nodes, [I64] * 4, [0] * 4, [5], predicate=7)


@pytest.mark.parametrize("rows,k", [(60_000_000, 1 << 19), (10_000_001, 1 << 12), (5_000_000, 1 << 20), (4097, 0)])
def test_c2_device_resident_bit_exact(ctx, rows, k):
    """BASELINE config 2 shape (scaled): every kept row, in order, equals the host recomputation."""
    d = {}
    for name in "abcd":
        kind, lo, span = GEN[name]
        d[name] = ctx.malloc(rows * 8 + 256)
        ctx.generate(d[name], rows, 0, 42, "abcd".index(name), kind, lo, span)
    out = ctx.malloc(rows * 8 + 256)
    prog = c2_program(ctx, k)
    kept = prog.run_sync([(d[c], None, capi.INT64) for c in "abcd"], rows, [(out, None, capi.INT64)])
    got = np.empty(kept, dtype=np.int64)
    if kept:
        ctx.d2h(got, out)
    # host recomputation in chunks (the 1B-row table does not fit host RAM of every box)
    pos = 0
    step = 8_000_000
    for first in range(0, rows, step):
        m = min(step, rows - first)
        a, b, c_, dd = (host_col(x, m, first) for x in "abcd")
        e = (a * b + c_)[dd < k]
        assert np.array_equal(got[pos:pos + len(e)], e), "mismatch in chunk starting at row %d" % first
        pos += len(e)
    assert pos == kept
    prog.close()
    for p in list(d.values()) + [out]:
        ctx.free(p)


def test_two_contexts_filter_concurrently(built):
    """Two contexts (streams) run Filter kernels at the same time, as the two lanes of the streaming cursor do.
    The kernel's CTAs wait on each other's kept-row counts, so each grid must be resident as a whole: the
    cooperative launch guarantees it whatever else shares the device. Forty launches per
    context from two host threads; every result must equal the host's."""
    import threading
    rows = 6_000_011
    k = 1 << 19
    want = None
    a, b, c_, dd = (host_col(x, rows, 0) for x in "abcd")
    want = (a * b + c_)[dd < k]
    errors = []

    def worker(tag):
        try:
            cx = capi.Context(0)
            d = {}
            for name in "abcd":
                kind, lo, span = GEN[name]
                d[name] = cx.malloc(rows * 8 + 256)
                cx.generate(d[name], rows, 0, 42, "abcd".index(name), kind, lo, span)
            out = cx.malloc(rows * 8 + 256)
            prog = c2_program(cx, k)
            for it in range(40):
                kept = prog.run_sync([(d[c], None, capi.INT64) for c in "abcd"], rows, [(out, None, capi.INT64)])
                if kept != len(want):
                    errors.append((tag, it, kept))
                    break
                if it % 13 == 0:
                    got = np.empty(kept, dtype=np.int64)
                    cx.d2h(got, out)
                    if not np.array_equal(got, want):
                        errors.append((tag, it, "values"))
                        break
            prog.close()
            for ptr in list(d.values()) + [out]:
                cx.free(ptr)
            cx.close()
        except Exception as e:   # noqa: BLE001
            errors.append((tag, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not any(t.is_alive() for t in threads), "a Filter kernel did not finish (co-residency deadlock)"
    assert not errors, errors


def test_filter_is_idempotent_and_order_preserving(ctx):
    rows = 20_000_000
    dcol = ctx.malloc(rows * 8 + 256)
    seq = ctx.malloc(rows * 8 + 256)
    ctx.generate(dcol, rows, 0, 42, 3, 0, 0, 1 << 20)
    ctx.h2d(seq, np.arange(rows, dtype=np.int64))
    n, I64, B = capi.node, capi.INT64, capi.BOOL
    nodes = [n(capi.OP_INPUT, I64, [0]), n(capi.OP_INPUT, I64, [1]), n(capi.OP_CONST, I64, [], i64=1 << 18), n(capi.OP_LT, B, [0, 2])]
    prog = capi.Program(ctx, nodes, [I64, I64], [0, 0], [0, 1], predicate=3)
    o1, o2 = ctx.malloc(rows * 8 + 256), ctx.malloc(rows * 8 + 256)
    k1 = prog.run_sync([(dcol, None, I64), (seq, None, I64)], rows, [(o1, None, I64), (o2, None, I64)])
    ids = np.empty(k1, dtype=np.int64)
    ctx.d2h(ids, o2)
    assert np.all(np.diff(ids) > 0)                       # input order kept
    p1, p2 = ctx.malloc(k1 * 8 + 256), ctx.malloc(k1 * 8 + 256)
    k2 = prog.run_sync([(o1, None, I64), (o2, None, I64)], k1, [(p1, None, I64), (p2, None, I64)])
    assert k2 == k1                                        # filtering the result again keeps everything
    ids2 = np.empty(k2, dtype=np.int64)
    ctx.d2h(ids2, p2)
    assert np.array_equal(ids, ids2)
    prog.close()
    for p in [dcol, seq, o1, o2, p1, p2]:
        ctx.free(p)


def _cols(items):
    arr = (capi.Column * max(1, len(items)))()
    for i, (d, n, t) in enumerate(items):
        arr[i].data, arr[i].nulls, arr[i].dtype = d, n, t
    return arr


def _group(ctx, specs_def, expected):
    specs = (capi.AggSpec * len(specs_def))()
    for i, (fn, inp, it, ot) in enumerate(specs_def):
        specs[i].fn, specs[i].input, specs[i].in_type, specs[i].out_type = fn, inp, it, ot
    g = C.c_void_p()
    kt, kn = (C.c_int32 * 1)(capi.INT64), (C.c_int32 * 1)(0)
    ctx.check(ctx.lib.ssb_group_create(ctx.h, 1, kt, kn, len(specs_def), specs, expected, C.byref(g)))
    return g


def _finalize(ctx, g, n_aggs, dtypes):
    n = C.c_int64()
    ko, ao = _cols([(0, None, 0)]), _cols([(0, None, 0)] * n_aggs)
    ctx.check(ctx.lib.ssb_group_finalize(g, C.byref(n), ko, ao))
    keys = np.empty(n.value, dtype=np.int64)
    ctx.d2h(keys, ko[0].data)
    outs = []
    for i, dt in enumerate(dtypes):
        a = np.empty(n.value, dtype=dt)
        ctx.d2h(a, ao[i].data)
        outs.append(a)
    order = np.argsort(keys)
    return keys[order], [a[order] for a in outs], (ko, ao, n.value)


@pytest.mark.parametrize("groups", [7, 5000, 1_000_000])
def test_group_chunked_updates_and_merge_equal_single_pass(ctx, groups):
    """C3 shape: aggregating two halves separately and merging the partial tables
    (ssb_group_merge: the multi-GPU exchange step) equals one pass, equals the host."""
    rows = 6_000_000
    k = ctx.malloc(rows * 8 + 256)
    v = ctx.malloc(rows * 8 + 256)
    ctx.generate(k, rows, 0, 42, 0, 1, 0, groups)
    ctx.generate(v, rows, 0, 42, 1, 2, 0, 0)
    spec = [(capi.AGG_SUM, 0, capi.DOUBLE, capi.DOUBLE), (capi.AGG_COUNT, -1, capi.INT64, capi.UINT64),
            (capi.AGG_MAX, 0, capi.DOUBLE, capi.DOUBLE)]
    dts = [np.float64, np.uint64, np.float64]
    # one pass
    g = _group(ctx, spec, 0)
    ctx.check(ctx.lib.ssb_group_update(g, _cols([(k, None, capi.INT64)]), _cols([(v, None, capi.DOUBLE)]), rows))
    k1, a1, _ = _finalize(ctx, g, 3, dts)
    # two chunked updates into one table
    g2 = _group(ctx, spec, 16)
    half = (rows // 2 // 32) * 32
    ctx.check(ctx.lib.ssb_group_update(g2, _cols([(k, None, capi.INT64)]), _cols([(v, None, capi.DOUBLE)]), half))
    ctx.check(ctx.lib.ssb_group_update(g2, _cols([(k + half * 8, None, capi.INT64)]), _cols([(v + half * 8, None, capi.DOUBLE)]), rows - half))
    k2, a2, _ = _finalize(ctx, g2, 3, dts)
    # two tables merged
    ga, gb = _group(ctx, spec, 0), _group(ctx, spec, 0)
    ctx.check(ctx.lib.ssb_group_update(ga, _cols([(k, None, capi.INT64)]), _cols([(v, None, capi.DOUBLE)]), half))
    ctx.check(ctx.lib.ssb_group_update(gb, _cols([(k + half * 8, None, capi.INT64)]), _cols([(v + half * 8, None, capi.DOUBLE)]), rows - half))
    _, _, (kob, aob, nb) = _finalize(ctx, gb, 3, dts)
    ctx.check(ctx.lib.ssb_group_merge(ga, nb, kob, aob))
    k3, a3, _ = _finalize(ctx, ga, 3, dts)
    # host
    hk = np.empty(rows, dtype=np.int64)
    hv = np.empty(rows, dtype=np.float64)
    ctx.lib.ssb_generate_host(hk.ctypes.data, rows, 0, 42, 0, 1, 0, groups)
    ctx.lib.ssb_generate_host(hv.ctypes.data, rows, 0, 42, 1, 2, 0, 0)
    uk, inv = np.unique(hk, return_inverse=True)
    hs = np.bincount(inv, weights=hv, minlength=len(uk))      # exactly summable payload: order-free
    hc = np.bincount(inv, minlength=len(uk)).astype(np.uint64)
    hm = np.full(len(uk), -np.inf)
    np.maximum.at(hm, inv, hv)
    for kk, aa in [(k1, a1), (k2, a2), (k3, a3)]:
        assert np.array_equal(kk, uk)
        assert np.array_equal(aa[0], hs) and np.array_equal(aa[1], hc) and np.array_equal(aa[2], hm)
    for h in [g, g2, ga, gb]:
        ctx.lib.ssb_group_destroy(h)
    ctx.free(k)
    ctx.free(v)


def test_sort_permutation_is_sorted_and_a_permutation(ctx):
    rows = 3_000_000
    k = ctx.malloc(rows * 8 + 256)
    ctx.generate(k, rows, 0, 42, 9, 0, -(1 << 40), 1 << 41)
    perm = ctx.malloc(rows * 8 + 256)
    for desc in (0, 1):
        ctx.check(ctx.lib.ssb_sort_permutation(ctx.h, 1, _cols([(k, None, capi.INT64)]), (C.c_int32 * 1)(desc), rows, perm))
        p = np.empty(rows, dtype=np.int64)
        ctx.d2h(p, perm)
        hk = np.empty(rows, dtype=np.int64)
        ctx.lib.ssb_generate_host(hk.ctypes.data, rows, 0, 42, 9, 0, -(1 << 40), 1 << 41)
        assert np.array_equal(np.sort(p), np.arange(rows))
        s = hk[p]
        assert np.all(np.diff(s) >= 0) if not desc else np.all(np.diff(s) <= 0)
        # stable: equal keys keep ascending row ids
        eq = np.diff(s) == 0
        assert np.all(np.diff(p)[eq] > 0)
    ctx.free(k)
    ctx.free(perm)


@pytest.mark.parametrize("rows,lo,span", [(1, 0, 16), (2, 0, 0), (4095, 0, 1 << 8), (4096, -5, 8), (4097, 0, 0),
                                          (1_000_001, 0, 1 << 20), (5_000_000, 0, 0), (20_000_000, -(1 << 62), 1 << 63)])
def test_sort_shapes_match_numpy_stable_argsort(ctx, rows, lo, span):
    """The one-sweep radix sort against numpy's stable argsort: partial tiles, keys whose high
    digits never vary (skipped passes), full 64-bit keys, duplicates (stability)."""
    k = ctx.malloc(rows * 8 + 256)
    perm = ctx.malloc(rows * 8 + 256)
    ctx.generate(k, rows, 0, 42, 11, 0, lo, span)
    hk = np.empty(rows, dtype=np.int64)
    ctx.lib.ssb_generate_host(hk.ctypes.data, rows, 0, 42, 11, 0, lo, span)
    ctx.check(ctx.lib.ssb_sort_permutation(ctx.h, 1, _cols([(k, None, capi.INT64)]), (C.c_int32 * 1)(0), rows, perm))
    p = np.empty(rows, dtype=np.int64)
    ctx.d2h(p, perm)
    assert np.array_equal(p, np.argsort(hk, kind="stable"))
    ctx.free(k)
    ctx.free(perm)


@pytest.mark.parametrize("rows,parts", [(1, 2), (1000, 1), (100_003, 2), (3_000_000, 8), (1_000_000, 200)])
def test_partition_rows_is_stable_and_sides_agree(ctx, rows, parts):
    """ssb_partition_rows: a permutation grouped by part with ascending row ids inside each
    part, counts that add up, and the same key lands in the same part whatever its row,
    column width (INT32 vs INT64) or side of the join."""
    k = ctx.malloc(rows * 8 + 256)
    ctx.generate(k, rows, 0, 42, 12, 1, -500, 1000 if rows > 1000 else 7)
    hk = np.empty(rows, dtype=np.int64)
    ctx.lib.ssb_generate_host(hk.ctypes.data, rows, 0, 42, 12, 1, -500, 1000 if rows > 1000 else 7)
    k32 = ctx.malloc(rows * 4 + 256)
    ctx.h2d(k32, hk.astype(np.int32))
    perm = ctx.malloc(rows * 8 + 256)
    got = []
    for ptr, dt in [(k, capi.INT64), (k32, capi.INT32)]:
        counts = (C.c_int64 * parts)()
        ctx.check(ctx.lib.ssb_partition_rows(ctx.h, 1, _cols([(ptr, None, dt)]), rows, parts, 0, perm, counts))
        p = np.empty(rows, dtype=np.int64)
        ctx.d2h(p, perm)
        cnt = np.array(list(counts))
        assert cnt.sum() == rows and np.array_equal(np.sort(p), np.arange(rows))
        part_of_key = {}
        off = 0
        for part, c in enumerate(cnt):
            seg = p[off:off + c]
            assert np.all(np.diff(seg) > 0)                      # stable inside a part
            for key in np.unique(hk[seg]):
                assert part_of_key.setdefault(int(key), part) == part   # a key lives in one part
            off += c
        got.append(part_of_key)
    assert got[0] == got[1]
    if parts >= 2 and rows >= 100_000:
        assert len(set(got[0].values())) == min(parts, len(set(got[0].values())))
        assert len(set(got[0].values())) > 1                      # the hash spreads keys
    for ptr in (k, k32, perm):
        ctx.free(ptr)


# ------------------------------------------------------------------------------------------------
# Fused Filter -> Compute -> GroupAggregate (ssb_group_update_program) against the unfused pipeline
# (ssb_program_run, then ssb_group_update) on the same inputs: identical groups and aggregates.
def _upload(ctx, arr, nulls=None):
    arr = np.ascontiguousarray(arr)
    ptr = ctx.malloc(arr.nbytes + 256)
    ctx.h2d(ptr, arr)
    nptr = None
    if nulls is not None:
        words = np.packbits(np.concatenate([nulls.astype(np.uint8), np.zeros((-len(nulls)) % 32 + 32, np.uint8)]), bitorder="little")
        nptr = ctx.malloc(words.nbytes + 256)
        ctx.h2d(nptr, words)
    return ptr, nptr


def _download_group(ctx, g, key_dts, agg_dts):
    n = C.c_int64()
    ko, ao = _cols([(0, None, 0)] * max(1, len(key_dts))), _cols([(0, None, 0)] * len(agg_dts))
    ctx.check(ctx.lib.ssb_group_finalize(g, C.byref(n), ko, ao))
    rows = n.value

    def col(c, dt):
        a = np.empty(rows, dtype=dt)
        if rows:
            ctx.d2h(a, c.data)
        isn = np.zeros(rows, dtype=bool)
        if c.nulls and rows:
            w = np.empty((rows + 31) // 32, dtype=np.uint32)
            ctx.d2h(w, c.nulls)
            isn = np.unpackbits(w.view(np.uint8), bitorder="little")[:rows].astype(bool)
        return a, isn
    keys = [col(ko[i], dt) for i, dt in enumerate(key_dts)]
    aggs = [col(ao[i], dt) for i, dt in enumerate(agg_dts)]
    # canonical order: by (null flags, key values)
    order = np.lexsort([x for k in reversed(keys) for x in (np.where(k[1], 0, k[0]), k[1])]) if keys else np.arange(rows)
    return [(k[0][order], k[1][order]) for k in keys], [(a[0][order], a[1][order]) for a in aggs]


def _same_groups(a, b):
    for (x, xn), (y, yn) in zip(a[0] + a[1], b[0] + b[1]):
        assert np.array_equal(xn, yn)
        assert np.array_equal(x[~xn], y[~yn])


NP_OF_T = {capi.INT64: np.int64, capi.DOUBLE: np.float64, capi.INT32: np.int32, capi.UINT64: np.uint64, capi.BOOL: np.uint8}


def _run_both(ctx, nodes, in_types, in_nullable, outputs, predicate, inputs, n_keys, key_types, key_nullable, aggs, agg_out_types):
    """inputs: [(values, nulls or None)]; aggs: [(fn, value index or -1, in_type, out_type, in_nullable)]."""
    rows = len(inputs[0][0])
    dev = [_upload(ctx, v, nl) for v, nl in inputs]
    incols = [(d, nl, t) for (d, nl), t in zip(dev, in_types)]
    prog = capi.Program(ctx, nodes, in_types, in_nullable, outputs, predicate=predicate)
    lib = ctx.lib

    def make_group():
        specs = (capi.AggSpec * len(aggs))()
        for i, (fn, inp, it, ot, inn) in enumerate(aggs):
            specs[i].fn, specs[i].input, specs[i].in_type, specs[i].out_type, specs[i].in_nullable = fn, inp, it, ot, inn
        g = C.c_void_p()
        kt = (C.c_int32 * max(1, n_keys))(*key_types)
        kn = (C.c_int32 * max(1, n_keys))(*key_nullable)
        ctx.check(lib.ssb_group_create(ctx.h, n_keys, kt, kn, len(aggs), specs, 0, C.byref(g)))
        return g
    # fused
    g1 = make_group()
    ctx.check(lib.ssb_group_update_program(g1, prog.h, _cols(incols), rows))
    key_dts = [NP_OF_T[t] for t in key_types]
    got = _download_group(ctx, g1, key_dts, [NP_OF_T[t] for t in agg_out_types])
    # unfused
    out_cols = []
    for j in range(len(outputs)):
        t = lib.ssb_program_output_type(prog.h, j)
        d = ctx.malloc(rows * 8 + 256)
        nl = ctx.malloc((rows // 32 + 2) * 4 + 256) if lib.ssb_program_output_nullable(prog.h, j) else None
        out_cols.append((d, nl, t))
    kept = prog.run_sync(incols, rows, out_cols)
    g2 = make_group()
    ctx.check(lib.ssb_group_update(g2, _cols(out_cols[:n_keys] or [(0, None, 0)]), _cols(out_cols[n_keys:] or [(0, None, 0)]), kept))
    want = _download_group(ctx, g2, key_dts, [NP_OF_T[t] for t in agg_out_types])
    _same_groups(got, want)
    for g in (g1, g2):
        lib.ssb_group_destroy(g)
    prog.close()
    for d, nl in dev:
        ctx.free(d)
        if nl:
            ctx.free(nl)
    for d, nl, _ in out_cols:
        ctx.free(d)
        if nl:
            ctx.free(nl)
    return want, kept


@pytest.fixture(params=["sink", "two_kernels", "jit"])
def sink_mode(request, monkeypatch):
    """ssb_group_update_program with the aggregation sink inside the expression kernel (the default for small
    inputs), with the two-kernel form (SSB200_GROUP_SINK=0) and with the kernel compiled at run time for the plan
    (SSB200_GROUP_JIT=1: csrc/jit.cu; by default only calls of 64M rows and more take it); the environment is
    read at every call."""
    monkeypatch.setenv("SSB200_GROUP_SINK", "0" if request.param == "two_kernels" else "1")
    monkeypatch.setenv("SSB200_GROUP_JIT", "1" if request.param == "jit" else "0")
    return request.param


@pytest.mark.parametrize("rows", [1, 1000, 600_000, 3_000_001])
def test_fused_q1_shape_equals_unfused(ctx, rows, sink_mode):
    rng = np.random.default_rng(rows)
    F64, I64, B = capi.DOUBLE, capi.INT64, capi.BOOL
    n = capi.node
    types = [F64, F64, F64, F64, I64, I64, I64]
    inputs = [(rng.integers(1, 51, rows).astype(np.float64), None), (rng.integers(100, 10000, rows) / 4.0, None),
              (rng.integers(0, 5, rows) / 16.0, None), (rng.integers(0, 5, rows) / 16.0, None),
              (rng.integers(0, 3, rows), None), (rng.integers(0, 2, rows), None), (rng.integers(0, 2500, rows), None)]
    nodes = [n(capi.OP_INPUT, t, [j]) for j, t in enumerate(types)]
    nodes += [n(capi.OP_CONST, F64, [], f64=1.0), n(capi.OP_SUB, F64, [7, 2]), n(capi.OP_MUL, F64, [1, 8]),
              n(capi.OP_ADD, F64, [7, 3]), n(capi.OP_MUL, F64, [9, 10]), n(capi.OP_CONST, I64, [], i64=2450),
              n(capi.OP_LE, B, [6, 12])]
    aggs = [(capi.AGG_SUM, i, F64, F64, 0) for i in range(5)] + [(capi.AGG_COUNT, -1, I64, capi.UINT64, 0)]
    want, kept = _run_both(ctx, nodes, types, [0] * 7, [4, 5, 0, 1, 2, 9, 11], 13, inputs, 2, [I64, I64], [0, 0], aggs,
                           [F64] * 5 + [capi.UINT64])
    assert int(want[1][5][0].sum()) == kept


@pytest.mark.parametrize("groups,rows", [(5, 50_000), (7, 400_000), (5000, 400_000)])
def test_fused_nullable_keys_inputs_and_overflow_equal_unfused(ctx, groups, rows, sink_mode):
    """NULL keys form a group, NULL predicate rows are dropped, NULL inputs do not count; with
    5000 groups the CTA-local entries overflow into the global table and later slices take the
    materialising path."""
    rng = np.random.default_rng(groups)
    F64, I64, I32, B = capi.DOUBLE, capi.INT64, capi.INT32, capi.BOOL
    n = capi.node
    types = [I64, I64, F64, I32]
    inputs = [(rng.integers(0, groups, rows), rng.random(rows) < 0.05), (rng.integers(-1000, 1000, rows), rng.random(rows) < 0.1),
              (rng.integers(0, 1 << 20, rows) / 8.0, rng.random(rows) < 0.2), (rng.integers(-50, 50, rows).astype(np.int32), None)]
    nodes = [n(capi.OP_INPUT, t, [j]) for j, t in enumerate(types)]                # 0..3
    nodes += [n(capi.OP_CONST, I64, [], i64=3), n(capi.OP_MUL, I64, [1, 4]),         # 5: b * 3 (nullable)
              n(capi.OP_CONST, I64, [], i64=-2500), n(capi.OP_GT, B, [5, 6]),         # 7: b*3 > -2500 (NULL drops the row)
              n(capi.OP_CAST, I64, [3]), n(capi.OP_ADD, I64, [5, 8])]                # 9: b*3 + d
    aggs = [(capi.AGG_SUM, 0, I64, I64, 1), (capi.AGG_MIN, 1, F64, F64, 1), (capi.AGG_MAX, 1, F64, F64, 1),
            (capi.AGG_COUNT, 1, F64, capi.UINT64, 1), (capi.AGG_COUNT, -1, I64, capi.UINT64, 0), (capi.AGG_SUM, 2, I32, I32, 0)]
    _run_both(ctx, nodes, types, [1, 1, 1, 0], [0, 9, 2, 3], 7, inputs, 1, [I64], [1], aggs,
              [I64, F64, F64, capi.UINT64, capi.UINT64, I32])


def test_fused_scalar_aggregate_equals_unfused(ctx, sink_mode):
    rows = 300_000
    rng = np.random.default_rng(3)
    I64, B = capi.INT64, capi.BOOL
    n = capi.node
    inputs = [(rng.integers(-10**6, 10**6, rows), None), (rng.integers(0, 100, rows), None)]
    nodes = [n(capi.OP_INPUT, I64, [0]), n(capi.OP_INPUT, I64, [1]), n(capi.OP_CONST, I64, [], i64=10), n(capi.OP_LT, B, [1, 2]),
             n(capi.OP_MUL, I64, [0, 1])]
    aggs = [(capi.AGG_SUM, 0, I64, I64, 0), (capi.AGG_MAX, 0, I64, I64, 0), (capi.AGG_COUNT, -1, I64, capi.UINT64, 0)]
    want, kept = _run_both(ctx, nodes, [I64, I64], [0, 0], [4], 3, inputs, 0, [], [], aggs, [I64, I64, capi.UINT64])
    a, b = inputs[0][0], inputs[1][0]
    m = b < 10
    assert want[1][0][0][0] == (a * b)[m].sum() and want[1][1][0][0] == (a * b)[m].max() and want[1][2][0][0] == m.sum() == kept


def test_fused_sink_overflow_and_table_growth_equal_unfused(ctx, sink_mode):
    """The aggregation sink of the expression kernel starts with four groups (per-thread
    accumulators), then the key domain explodes: rows overflow into the global table, the table
    fills up, the overflowing rows are deferred, the table grows and the rows are replayed."""
    rows = 500_000
    I64, B = capi.INT64, capi.BOOL
    n = capi.node
    i = np.arange(rows, dtype=np.int64)
    key = np.where(i < 200_000, i % 4, i)
    val = (i * 7919) % 1000 - 500
    inputs = [(key, None), (val, None)]
    nodes = [n(capi.OP_INPUT, I64, [0]), n(capi.OP_INPUT, I64, [1]), n(capi.OP_CONST, I64, [], i64=400), n(capi.OP_LT, B, [1, 2]),
             n(capi.OP_CONST, I64, [], i64=2), n(capi.OP_MUL, I64, [1, 4])]
    aggs = [(capi.AGG_SUM, 0, I64, I64, 0), (capi.AGG_MIN, 0, I64, I64, 0), (capi.AGG_COUNT, -1, I64, capi.UINT64, 0)]
    want, kept = _run_both(ctx, nodes, [I64, I64], [0, 0], [0, 5], 3, inputs, 1, [I64], [0], aggs, [I64, I64, capi.UINT64])
    assert int(want[1][2][0].sum()) == kept == int((val < 400).sum())


# ------------------------------------------------------------------------------------------------
# BASELINE.json's full sizes (1B rows): size-independent checks against a chunked host recomputation
# of the same counter-based columns (ssb_generate_host is bit-identical to the device generator).
def _host_chunks(total, chunk):
    first = 0
    while first < total:
        n = min(chunk, total - first)
        yield first, n
        first += n


def _gen_host(ctx, n, first, stream, kind, lo, span, dtype=np.int64):
    out = np.empty(n, dtype=dtype)
    ctx.lib.ssb_generate_host(out.ctypes.data, n, first, 42, stream, kind, lo, span)
    return out


def test_c2_full_size_count_and_checksum(ctx):
    """Config 2 at 1B rows: the number of kept rows and the wrapping 64-bit sum of the kept
    e = a*b+c values (a checksum of the whole ordered output) equal the host's, and the output is
    the host's output on three windows spread over the result (order check)."""
    rows = 1_000_000_000
    k = 1 << 19
    d = {}
    for name in "abcd":
        kind, lo, span = GEN[name]
        d[name] = ctx.malloc(rows * 8 + 256)
        ctx.generate(d[name], rows, 0, 42, "abcd".index(name), kind, lo, span)
    out = ctx.malloc(rows * 8 + 256)
    prog = c2_program(ctx, k)
    kept = prog.run_sync([(d[c], None, capi.INT64) for c in "abcd"], rows, [(out, None, capi.INT64)])
    # device-side checksum: ScalarAggregate SUM (wrapping) + COUNT over the kept values
    specs = (capi.AggSpec * 2)()
    specs[0].fn, specs[0].input, specs[0].in_type, specs[0].out_type = capi.AGG_SUM, 0, capi.INT64, capi.INT64
    specs[1].fn, specs[1].input, specs[1].in_type, specs[1].out_type = capi.AGG_COUNT, -1, capi.INT64, capi.UINT64
    g = C.c_void_p()
    dummy = (C.c_int32 * 1)(0)
    ctx.check(ctx.lib.ssb_group_create(ctx.h, 0, dummy, dummy, 2, specs, 0, C.byref(g)))
    ctx.check(ctx.lib.ssb_group_update(g, _cols([(0, None, 0)]), _cols([(out, None, capi.INT64)]), kept))
    n = C.c_int64()
    ko, ao = _cols([(0, None, 0)]), _cols([(0, None, 0), (0, None, 0)])
    ctx.check(ctx.lib.ssb_group_finalize(g, C.byref(n), ko, ao))
    dev_sum, dev_cnt = np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.uint64)
    ctx.d2h(dev_sum, ao[0].data)
    ctx.d2h(dev_cnt, ao[1].data)
    ctx.lib.ssb_group_destroy(g)
    # host: chunked recomputation
    host_cnt, host_sum = 0, np.int64(0)
    windows = {}
    want_windows = [0, kept // 2, kept - 1000]
    with np.errstate(over="ignore"):
        for first, cn in _host_chunks(rows, 50_000_000):
            cols = {c: _gen_host(ctx, cn, first, "abcd".index(c), *GEN[c]) for c in "abcd"}
            m = cols["d"] < k
            e = (cols["a"] * cols["b"] + cols["c"])[m]
            for w0 in want_windows:          # output positions [w0, w0 + 1000) that fall into this chunk
                lo_, hi_ = max(w0, host_cnt), min(w0 + 1000, host_cnt + len(e))
                if lo_ < hi_:
                    windows.setdefault(w0, []).append(e[lo_ - host_cnt:hi_ - host_cnt])
            host_cnt += int(m.sum())
            host_sum = host_sum + e.sum(dtype=np.int64)
    assert kept == host_cnt == int(dev_cnt[0])
    assert int(dev_sum[0]) == int(host_sum)
    for w0 in want_windows:
        got = np.empty(1000, dtype=np.int64)
        ctx.d2h(got, out + w0 * 8)
        assert np.array_equal(got, np.concatenate(windows[w0]))
    prog.close()
    for name in "abcd":
        ctx.free(d[name])
    ctx.free(out)


def test_c3_full_size_equals_host(ctx):
    """Config 3 at 1B rows, 1M INT64 keys, SUM(DOUBLE) + COUNT(*): every group's sum (exactly
    summable payload: order-free) and count equal the host's bincount over the same rows."""
    rows, groups = 1_000_000_000, 1_000_000
    k = ctx.malloc(rows * 8 + 256)
    v = ctx.malloc(rows * 8 + 256)
    ctx.generate(k, rows, 0, 42, 20, 1, 0, groups)
    ctx.generate(v, rows, 0, 42, 21, 2, 0, 0)
    spec = [(capi.AGG_SUM, 0, capi.DOUBLE, capi.DOUBLE), (capi.AGG_COUNT, -1, capi.INT64, capi.UINT64)]
    g = _group(ctx, spec, groups)
    ctx.check(ctx.lib.ssb_group_update(g, _cols([(k, None, capi.INT64)]), _cols([(v, None, capi.DOUBLE)]), rows))
    keys, aggs, _ = _finalize(ctx, g, 2, [np.float64, np.uint64])
    ctx.lib.ssb_group_destroy(g)
    ctx.free(k)
    ctx.free(v)
    hs, hc = np.zeros(groups), np.zeros(groups, dtype=np.int64)
    for first, cn in _host_chunks(rows, 50_000_000):
        hk = _gen_host(ctx, cn, first, 20, 1, 0, groups)
        hv = _gen_host(ctx, cn, first, 21, 2, 0, 0, dtype=np.float64)
        hs += np.bincount(hk, weights=hv, minlength=groups)
        hc += np.bincount(hk, minlength=groups)
    assert np.array_equal(keys, np.arange(groups))
    assert np.array_equal(aggs[1], hc.astype(np.uint64))
    assert np.array_equal(aggs[0], hs)


def test_c4_full_size_join_pairs_in_order(ctx):
    """Config 4 on one GPU: 1B probe rows against a 100M-row build side whose key column is a
    permutation of [0, 1e8) (pk(i) = i * m mod B). Every probe row matches exactly once, the pairs
    come in probe order, and the matched build row is the one the host computes through the
    modular inverse of the permutation (checked on windows spread over the result)."""
    probe, build = 1_000_000_000, 100_000_000
    mult = 1_000_000_007
    pk = ctx.malloc(build * 8 + 256)
    fk = ctx.malloc(probe * 8 + 256)
    ctx.generate(pk, build, 0, 42, 40, 4, mult, build)
    ctx.generate(fk, probe, 0, 42, 42, 1, 0, build)
    j = C.c_void_p()
    ctx.check(ctx.lib.ssb_join_build(ctx.h, 1, _cols([(pk, None, capi.INT64)]), build, 1, C.byref(j)))
    n, pl, pr = C.c_int64(), C.c_void_p(), C.c_void_p()
    ctx.check(ctx.lib.ssb_join_probe(j, _cols([(fk, None, capi.INT64)]), probe, 0, C.byref(n), C.byref(pl), C.byref(pr)))
    assert n.value == probe
    inv = pow(mult, -1, build)
    for w0 in [0, 123_456_789, probe // 2, probe - 4096]:
        li, ri = np.empty(4096, dtype=np.int64), np.empty(4096, dtype=np.int64)
        ctx.d2h(li, pl.value + w0 * 8)
        ctx.d2h(ri, pr.value + w0 * 8)
        assert np.array_equal(li, np.arange(w0, w0 + 4096))                       # probe order, one pair per row
        hfk = _gen_host(ctx, 4096, w0, 42, 1, 0, build)
        want = np.array([(int(x) * inv) % build for x in hfk], dtype=np.int64)    # the build row holding that key
        assert np.array_equal(ri, want)
    ctx.lib.ssb_join_destroy(j)
    ctx.free(pk)
    ctx.free(fk)


@pytest.mark.parametrize("compact", [0, 0x100])
@pytest.mark.parametrize("join_type", [0, 1])
def test_probe_materialize_equals_probe_plus_gather(ctx, join_type, compact):
    """ssb_join_probe_materialize (UNIQUE keys: the probe writes the result columns at their final positions) against
    ssb_join_probe + ssb_gather and against numpy: INT64 / INT32 / DOUBLE / BOOL columns, NULL probe keys, probe keys
    without a build row, both join types, regular and compact tables."""
    rng = np.random.default_rng(17 + join_type)
    nb, npr = 70_001, 300_007
    pk = rng.permutation(nb * 2)[:nb].astype(np.int64)
    pay = rng.integers(-2**62, 2**62, nb)
    flag = rng.integers(0, 2, nb).astype(np.bool_)
    fk = rng.integers(0, nb * 2, npr)
    fk_null = rng.random(npr) < 0.05
    lv = rng.integers(0, 1000, npr).astype(np.int32)
    ld = rng.random(npr)
    d_pk, _ = _upload(ctx, pk)
    d_pay, _ = _upload(ctx, pay)
    d_flag, _ = _upload(ctx, flag)
    d_fk, d_fkn = _upload(ctx, fk, fk_null)
    d_lv, _ = _upload(ctx, lv)
    d_ld, _ = _upload(ctx, ld)
    j = C.c_void_p()
    ctx.check(ctx.lib.ssb_join_build(ctx.h, 1, _cols([(d_pk, None, capi.INT64)]), nb, 1 | compact, C.byref(j)))
    outs = [(ctx.malloc(npr * 8 + 256), np.int32, capi.INT32), (ctx.malloc(npr * 8 + 256), np.float64, capi.DOUBLE),
            (ctx.malloc(npr * 8 + 256), np.int64, capi.INT64), (ctx.malloc(npr * 8 + 256), np.bool_, capi.BOOL)]
    matched = ctx.malloc(npr + 256)
    n = C.c_int64()
    ctx.check(ctx.lib.ssb_join_probe_materialize(j, _cols([(d_fk, d_fkn, capi.INT64)]), npr, join_type,
                                                 2, _cols([(d_lv, None, capi.INT32), (d_ld, None, capi.DOUBLE)]),
                                                 2, _cols([(d_pay, None, capi.INT64), (d_flag, None, capi.BOOL)]),
                                                 _cols([(p, None, dt) for p, _, dt in outs]), matched, C.byref(n)))
    row_of = {int(k): i for i, k in enumerate(pk)}
    hit = np.array([(-1 if fk_null[i] else row_of.get(int(fk[i]), -1)) for i in range(npr)])
    keep = np.arange(npr) if join_type == 1 else np.nonzero(hit >= 0)[0]
    assert n.value == len(keep)
    got = []
    for p, npdt, _ in outs:
        a = np.empty(n.value, dtype=npdt)
        ctx.d2h(a, p)
        got.append(a)
    h = hit[keep]
    assert np.array_equal(got[0], lv[keep]) and np.array_equal(got[1], ld[keep])
    assert np.array_equal(got[2], np.where(h >= 0, pay[np.maximum(h, 0)], 0))
    assert np.array_equal(got[3], np.where(h >= 0, flag[np.maximum(h, 0)], False))
    if join_type == 1:
        m = np.empty(npr, dtype=np.uint8)
        ctx.d2h(m, matched)
        assert np.array_equal(m.astype(bool), h >= 0)
    # the two-step form gives the same pairs
    n2, pl, pr = C.c_int64(), C.c_void_p(), C.c_void_p()
    ctx.check(ctx.lib.ssb_join_probe(j, _cols([(d_fk, d_fkn, capi.INT64)]), npr, join_type, C.byref(n2), C.byref(pl), C.byref(pr)))
    assert n2.value == n.value
    li, ri = np.empty(n2.value, dtype=np.int64), np.empty(n2.value, dtype=np.int64)
    ctx.d2h(li, pl)
    ctx.d2h(ri, pr)
    assert np.array_equal(li, keep) and np.array_equal(ri, h)
    ctx.lib.ssb_join_destroy(j)


@pytest.mark.parametrize("n_parts", [2, 5])
@pytest.mark.parametrize("join_type", [0, 1])
def test_attached_table_parts_equal_host(ctx, join_type, n_parts):
    """The replicated form of the sharded join on ONE GPU (what csrc/shard_join.cu does over NCCL): the build keys are
    hash-partitioned (ssb_partition_rows), every part gets its own compact table, the tables are attached as one index
    (ssb_join_attach_parts) and probed with the pairs-only and the materialising probe; NULL keys on both sides, keys
    without a build row. rhs rows are reported in part order (row_offsets[part] + row inside the part)."""
    rng = np.random.default_rng(n_parts * 10 + join_type)
    nb, npr = 50_000, 200_003
    pk = rng.permutation(nb * 50)[:nb].astype(np.int64) - 1_000_000          # sparse: the slot tables, not the dense index
    pk_null = rng.random(nb) < 0.03
    pay = rng.integers(-2**62, 2**62, nb)
    fk = rng.integers(-1_000_100, nb * 50 - 999_900, npr)
    fk_null = rng.random(npr) < 0.05
    lv = rng.integers(0, 1000, npr)
    d_pk, d_pkn = _upload(ctx, pk, pk_null)
    d_pay, _ = _upload(ctx, pay)
    d_fk, d_fkn = _upload(ctx, fk, fk_null)
    d_lv, _ = _upload(ctx, lv)
    lib = ctx.lib
    # partition the build rows: parts 0 .. n_parts-1 by key hash, NULL keys set aside in part n_parts
    d_perm = ctx.malloc(nb * 8 + 256)
    counts = (C.c_int64 * (n_parts + 1))()
    ctx.check(lib.ssb_partition_rows(ctx.h, 1, _cols([(d_pk, d_pkn, capi.INT64)]), nb, n_parts, n_parts, d_perm, counts))
    perm = np.empty(nb, dtype=np.int64)
    ctx.d2h(perm, d_perm)
    sent = sum(counts[p] for p in range(n_parts))
    assert counts[n_parts] == int(pk_null.sum()) and sent == nb - int(pk_null.sum())
    d_pk_parts = ctx.malloc(nb * 8 + 256)
    d_pay_parts = ctx.malloc(nb * 8 + 256)
    ctx.check(lib.ssb_gather(ctx.h, _cols([(d_pk, None, capi.INT64)]), d_perm, sent, _cols([(d_pk_parts, None, capi.INT64)])))
    ctx.check(lib.ssb_gather(ctx.h, _cols([(d_pay, None, capi.INT64)]), d_perm, sent, _cols([(d_pay_parts, None, capi.INT64)])))
    joins, slots, caps, offs = [], (C.c_void_p * n_parts)(), (C.c_int64 * n_parts)(), (C.c_int64 * n_parts)()
    at = 0
    for p in range(n_parts):
        j = C.c_void_p()
        ctx.check(lib.ssb_join_build(ctx.h, 1, _cols([(d_pk_parts + at * 8, None, capi.INT64)]), counts[p], 1 | 0x100, C.byref(j)))
        s_, c_ = C.c_void_p(), C.c_int64()
        ctx.check(lib.ssb_join_table(j, C.byref(s_), C.byref(c_)))
        joins.append(j)
        slots[p], caps[p], offs[p] = s_.value, c_.value, at
        at += counts[p]
    idx = C.c_void_p()
    ctx.check(lib.ssb_join_attach_parts(ctx.h, capi.INT64, n_parts, slots, caps, offs, C.byref(idx)))
    # expected: build rows renumbered in part order
    pos_in_parts = {int(pk[r]): i for i, r in enumerate(perm[:sent])}
    hit = np.array([(-1 if fk_null[i] else pos_in_parts.get(int(fk[i]), -1)) for i in range(npr)])
    keep = np.arange(npr) if join_type == 1 else np.nonzero(hit >= 0)[0]
    n, pl, pr = C.c_int64(), C.c_void_p(), C.c_void_p()
    ctx.check(lib.ssb_join_probe(idx, _cols([(d_fk, d_fkn, capi.INT64)]), npr, join_type, C.byref(n), C.byref(pl), C.byref(pr)))
    assert n.value == len(keep)
    li, ri = np.empty(n.value, dtype=np.int64), np.empty(n.value, dtype=np.int64)
    ctx.d2h(li, pl)
    ctx.d2h(ri, pr)
    assert np.array_equal(li, keep) and np.array_equal(ri, hit[keep])
    o_lv, o_pay = ctx.malloc(npr * 8 + 256), ctx.malloc(npr * 8 + 256)
    d_match = ctx.malloc(npr + 256)
    m = C.c_int64()
    ctx.check(lib.ssb_join_probe_materialize(idx, _cols([(d_fk, d_fkn, capi.INT64)]), npr, join_type, 1, _cols([(d_lv, None, capi.INT64)]),
                                             1, _cols([(d_pay_parts, None, capi.INT64)]),
                                             _cols([(o_lv, None, capi.INT64), (o_pay, None, capi.INT64)]), d_match, C.byref(m)))
    assert m.value == len(keep)
    got_lv, got_pay = np.empty(m.value, dtype=np.int64), np.empty(m.value, dtype=np.int64)
    ctx.d2h(got_lv, o_lv)
    ctx.d2h(got_pay, o_pay)
    h = hit[keep]
    pay_parts = pay[perm[:sent]]
    assert np.array_equal(got_lv, lv[keep])
    assert np.array_equal(got_pay, np.where(h >= 0, pay_parts[np.maximum(h, 0)], 0))
    lib.ssb_join_destroy(idx)
    for j in joins:
        lib.ssb_join_destroy(j)
    for d in (d_pk, d_pkn, d_pay, d_fk, d_fkn, d_lv, d_perm, d_pk_parts, d_pay_parts, o_lv, o_pay, d_match):
        ctx.free(d)


@pytest.mark.parametrize("build_dtype,probe_dtype", [(np.int64, np.int64), (np.int32, np.int64), (np.uint32, np.int32)])
@pytest.mark.parametrize("lo,spread", [(0, 2), (-50_000, 3), (2**31 - 120_000, 1), (-2**40, 2), (0, 1000)])
def test_join_dense_integer_keys_equal_host(ctx, lo, spread, build_dtype, probe_dtype):
    """UNIQUE joins on one integer key column whose values span at most 4 x rows use a direct index
    (dense_rows[key - min] = build row) instead of the slot table; spread 1000 keeps the slot table. Against numpy:
    negative and large bases, build and probe keys of different integer types (compared by value), NULL keys on both
    sides, probe keys below / above / inside the range without a build row, duplicate build keys (the smallest row is
    the head, as in the slot table), INNER and LEFT_OUTER."""
    nb, npr = 60_000, 250_003
    narrow = np.dtype(build_dtype).itemsize == 4 or np.dtype(probe_dtype).itemsize == 4
    unsigned = np.dtype(build_dtype).kind == "u" or np.dtype(probe_dtype).kind == "u"
    if (narrow and (lo - 500 < -2**31 or lo + nb * spread + 500 >= 2**31)) or (unsigned and lo < 0):
        pytest.skip("key range outside the 32-bit type")
    rng = np.random.default_rng(abs(lo) % 1000 + spread)
    pk = (rng.permutation(nb * spread)[:nb] + lo).astype(np.int64)
    pk[1000:1200] = pk[:200]                                 # duplicates: rows 0..199 stay the heads
    pk_null = rng.random(nb) < 0.02
    fk = (rng.integers(-500, nb * spread + 500, npr) + lo).astype(np.int64)
    fk_null = rng.random(npr) < 0.05
    lim = np.iinfo(probe_dtype)
    fk = np.clip(fk, lim.min, lim.max)
    d_pk, d_pkn = _upload(ctx, pk.astype(build_dtype), pk_null)
    d_fk, d_fkn = _upload(ctx, fk.astype(probe_dtype), fk_null)
    tb = {np.int64: capi.INT64, np.int32: capi.INT32, np.uint32: capi.UINT32}
    row_of = {}
    for i in range(nb - 1, -1, -1):
        if not pk_null[i]:
            row_of[int(pk[i])] = i
    hit = np.array([(-1 if fk_null[i] else row_of.get(int(fk[i]), -1)) for i in range(npr)])
    j = C.c_void_p()
    ctx.check(ctx.lib.ssb_join_build(ctx.h, 1, _cols([(d_pk, d_pkn, tb[build_dtype])]), nb, 1, C.byref(j)))
    for join_type in (0, 1):
        keep = np.arange(npr) if join_type == 1 else np.nonzero(hit >= 0)[0]
        n, pl, pr = C.c_int64(), C.c_void_p(), C.c_void_p()
        ctx.check(ctx.lib.ssb_join_probe(j, _cols([(d_fk, d_fkn, tb[probe_dtype])]), npr, join_type, C.byref(n), C.byref(pl), C.byref(pr)))
        assert n.value == len(keep)
        li, ri = np.empty(n.value, dtype=np.int64), np.empty(n.value, dtype=np.int64)
        ctx.d2h(li, pl)
        ctx.d2h(ri, pr)
        assert np.array_equal(li, keep) and np.array_equal(ri, hit[keep])
    ctx.lib.ssb_join_destroy(j)
    for d in (d_pk, d_pkn, d_fk, d_fkn):
        ctx.free(d)


@pytest.mark.parametrize("with_count", [True, False])
@pytest.mark.parametrize("lo", [0, -700_000, 2**62])
def test_group_dense_keys_equal_host(ctx, lo, with_count, monkeypatch):
    """Dense integer keys (slot = key - lo, two REDs per row instead of three L2 accesses): 600k distinct keys in a
    narrow range starting at `lo`, a sprinkle of far outliers (they take the general table) and keys that first appear
    after the range was chosen; SUM(DOUBLE) + SUM(INT64) with and without a COUNT(*) (without one a hit counter marks
    the touched slots); chunked updates and a merge of two tables. Equal to the host and to the general table
    (SSB200_GROUP_DENSE=0 is read once per process, so the general table is exercised through MIN, which the dense
    path does not take)."""
    rng = np.random.default_rng(11)
    rows = 3_000_000
    keys = rng.integers(0, 600_000, rows) + lo
    keys[:200_000] = rng.integers(100_000, 500_000, 200_000) + lo      # the first rows see only the middle of the range
    out = rng.integers(0, rows, 5_000)
    keys[out] = rng.integers(-2**62, 2**62, 5_000)                       # far outside
    keys[rng.integers(200_000, rows, 2_000)] = lo + 600_000 + 200_000    # beyond the margin
    keys = keys.astype(np.int64)
    vd = rng.integers(-2**20, 2**20, rows) / 1024.0
    vi = rng.integers(-2**40, 2**40, rows)
    d_k, _ = _upload(ctx, keys)
    d_vd, _ = _upload(ctx, vd)
    d_vi, _ = _upload(ctx, vi)
    spec = [(capi.AGG_SUM, 0, capi.DOUBLE, capi.DOUBLE), (capi.AGG_SUM, 1, capi.INT64, capi.INT64)]
    dts = [np.float64, np.int64]
    if with_count:
        spec.append((capi.AGG_COUNT, -1, capi.INT64, capi.UINT64))
        dts.append(np.uint64)
    kcol = lambda off=0: _cols([(d_k + off * 8, None, capi.INT64)])                                   # noqa: E731
    vcol = lambda off=0: _cols([(d_vd + off * 8, None, capi.DOUBLE), (d_vi + off * 8, None, capi.INT64)])   # noqa: E731
    g = _group(ctx, spec, 0)
    before = ctx.launches()
    ctx.check(ctx.lib.ssb_group_update(g, kcol(), vcol(), rows))
    k1, a1, _ = _finalize(ctx, g, len(spec), dts)
    # chunks + merge
    ga, gb = _group(ctx, spec, 0), _group(ctx, spec, 0)
    half = (rows // 2 // 32) * 32
    third = (half // 3 // 32) * 32
    ctx.check(ctx.lib.ssb_group_update(ga, kcol(), vcol(), third))
    ctx.check(ctx.lib.ssb_group_update(ga, kcol(third), vcol(third), half - third))
    ctx.check(ctx.lib.ssb_group_update(gb, kcol(half), vcol(half), rows - half))
    _, _, (kob, aob, nb) = _finalize(ctx, gb, len(spec), dts)
    ctx.check(ctx.lib.ssb_group_merge(ga, nb, kob, aob))
    k2, a2, _ = _finalize(ctx, ga, len(spec), dts)
    uk, inv = np.unique(keys, return_inverse=True)
    hs = np.bincount(inv, weights=vd, minlength=len(uk))
    hi = np.zeros(len(uk), dtype=np.int64)
    np.add.at(hi, inv, vi)
    want = [hs, hi] + ([np.bincount(inv, minlength=len(uk)).astype(np.uint64)] if with_count else [])
    for kk, aa in [(k1, a1), (k2, a2)]:
        assert np.array_equal(kk, uk)
        for got, w in zip(aa, want):
            assert np.array_equal(got, w)
    assert ctx.launches() > before
    for h in [g, ga, gb]:
        ctx.lib.ssb_group_destroy(h)
