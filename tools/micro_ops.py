"""Device-resident micro-benchmarks of group-by / join / sort through the C ABI."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
from supersonic_b200 import capi

ctx = capi.Context(0)
lib = ctx.lib
ctx.enable_timing(True)


def cols(items):
    arr = (capi.Column * max(1, len(items)))()
    for i, (d, n, t) in enumerate(items):
        arr[i].data, arr[i].nulls, arr[i].dtype = d, n, t
    return arr


def group_bench(rows, groups, label):
    k = ctx.malloc(rows * 8 + 256)
    v = ctx.malloc(rows * 8 + 256)
    ctx.generate(k, rows, 0, 42, 0, 1, 0, groups)
    ctx.generate(v, rows, 0, 42, 1, 2, 0, 0)
    specs = (capi.AggSpec * 2)()
    specs[0].fn, specs[0].input, specs[0].in_type, specs[0].out_type = capi.AGG_SUM, 0, capi.DOUBLE, capi.DOUBLE
    specs[1].fn, specs[1].input, specs[1].in_type, specs[1].out_type = capi.AGG_COUNT, -1, capi.INT64, capi.UINT64
    kt = (C.c_int32 * 1)(capi.INT64)
    kn = (C.c_int32 * 1)(0)
    best = None
    for it in range(3):
        g = C.c_void_p()
        ctx.check(lib.ssb_group_create(ctx.h, 1, kt, kn, 2, specs, groups, C.byref(g)))
        ctx.sync()
        ctx.timer_start()
        ctx.check(lib.ssb_group_update(g, cols([(k, None, capi.INT64)]), cols([(v, None, capi.DOUBLE)]), rows))
        ms = ctx.timer_stop()
        n = C.c_int64()
        ko, ao = cols([(0, None, 0)]), cols([(0, None, 0), (0, None, 0)])
        ctx.check(lib.ssb_group_finalize(g, C.byref(n), ko, ao))
        lib.ssb_group_destroy(g)
        best = ms if best is None else min(best, ms)
    print("%-34s rows=%d groups=%d  %.3f ms  %.1f Grows/s  %.1f GB/s (16 B/row)" % (label, rows, n.value, best, rows / best / 1e6, rows * 16 / best / 1e6))
    sys.stdout.flush()
    ctx.free(k); ctx.free(v)


def sort_bench(rows, span=0, label="full 64-bit keys"):
    k = ctx.malloc(rows * 8 + 256)
    ctx.generate(k, rows, 0, 42, 0, 0, 0, span)
    perm = ctx.malloc(rows * 8 + 256)
    desc = (C.c_int32 * 1)(0)
    best = None
    for _ in range(3):
        ctx.sync()
        ctx.timer_start()
        ctx.check(lib.ssb_sort_permutation(ctx.h, 1, cols([(k, None, capi.INT64)]), desc, rows, perm))
        ms = ctx.timer_stop()
        best = ms if best is None else min(best, ms)
    print("sort INT64 (%s) rows=%d  %.3f ms  %.2f Grows/s" % (label, rows, best, rows / best / 1e6)); sys.stdout.flush()
    ctx.free(k); ctx.free(perm)


def partition_bench(rows, parts):
    k = ctx.malloc(rows * 8 + 256)
    ctx.generate(k, rows, 0, 42, 3, 0, 0, 0)
    perm = ctx.malloc(rows * 8 + 256)
    counts = (C.c_int64 * parts)()
    best = None
    for _ in range(3):
        ctx.sync()
        ctx.timer_start()
        ctx.check(lib.ssb_partition_rows(ctx.h, 1, cols([(k, None, capi.INT64)]), rows, parts, 0, perm, counts))
        ms = ctx.timer_stop()
        best = ms if best is None else min(best, ms)
    print("partition rows=%d parts=%d  %.3f ms  %.2f Grows/s" % (rows, parts, best, rows / best / 1e6)); sys.stdout.flush()
    ctx.free(k); ctx.free(perm)


def join_bench(build, probe, uniq=1):
    pk = ctx.malloc(build * 8 + 256)
    fk = ctx.malloc(probe * 8 + 256)
    # pk: a permutation-like injective map (odd multiplier mod 2^k keeps uniqueness) is not needed
    # for timing: use i*3 (unique), fk uniform over [0, 3*build)
    h = np.arange(build, dtype=np.int64) * 3
    ctx.h2d(pk, h)
    ctx.generate(fk, probe, 0, 42, 5, 1, 0, build * 3)
    ctx.sync()
    bms = pms = None
    for _ in range(3):
        j = C.c_void_p()
        ctx.timer_start()
        ctx.check(lib.ssb_join_build(ctx.h, 1, cols([(pk, None, capi.INT64)]), build, uniq, C.byref(j)))
        b = ctx.timer_stop()
        n = C.c_int64(); l = C.c_void_p(); r = C.c_void_p()
        ctx.timer_start()
        ctx.check(lib.ssb_join_probe(j, cols([(fk, None, capi.INT64)]), probe, 0, C.byref(n), C.byref(l), C.byref(r)))
        p_ = ctx.timer_stop()
        bms = b if bms is None else min(bms, b)
        pms = p_ if pms is None else min(pms, p_)
        lib.ssb_join_destroy(j)
    print("join %s build=%d %.3f ms; probe=%d %.3f ms (%.2f Grows/s) pairs=%d" % ("UNIQUE" if uniq else "NOT_UNIQUE", build, bms, probe, pms, probe / pms / 1e6, n.value)); sys.stdout.flush()
    ctx.free(pk); ctx.free(fk)


if __name__ == "__main__":
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    group_bench(scale, 1_000_000, "C3 group-by SUM(double)+COUNT")
    group_bench(scale, 1000, "group-by 1000 groups")
    group_bench(scale, 6, "group-by 6 groups")
    sort_bench(scale // 4)
    sort_bench(scale // 4, 1 << 24, "24-bit keys")
    partition_bench(scale, 8)
    join_bench(scale // 10, scale)
    join_bench(scale // 10, scale, 0)
