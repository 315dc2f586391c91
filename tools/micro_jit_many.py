"""Fused Filter -> Compute -> GroupAggregate with MANY groups (sparse INT64 keys): the run-time compiled kernel without
CTA-local entries (every row to the global table) against the materialising slices.
usage: python tools/micro_jit_many.py [rows] [groups]"""
import ctypes as C
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from supersonic_b200 import capi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
ctx = capi.Context(0)
lib = ctx.lib
I64, F64, B = capi.INT64, capi.DOUBLE, capi.BOOL
gens = [(1, 0, groups), (5, 1600, 160000), (5, 0, 2), (1, 0, 2500)]   # key index, price, disc, ship
types = [I64, F64, F64, I64]
d_in = []
for j, (kind, lo, span) in enumerate(gens):
    p = ctx.malloc(rows * 8 + 256)
    ctx.generate(p, rows, 0, 7, 60 + j, kind, lo, span)
    d_in.append(p)
n = capi.node
nodes = [n(capi.OP_INPUT, t, [j]) for j, t in enumerate(types)]                   # 0..3
nodes += [n(capi.OP_CONST, I64, [], i64=7919), n(capi.OP_MUL, I64, [0, 4]),         # 5: sparse key
          n(capi.OP_CONST, F64, [], f64=1.0), n(capi.OP_SUB, F64, [6, 2]), n(capi.OP_MUL, F64, [1, 7]),   # 8: disc_price
          n(capi.OP_CONST, I64, [], i64=2450), n(capi.OP_LE, B, [3, 9])]            # 10: ship <= D
prog = capi.Program(ctx, nodes, types, [0] * 4, [5, 8], predicate=10)
specs = (capi.AggSpec * 2)()
specs[0].fn, specs[0].input, specs[0].in_type, specs[0].out_type = capi.AGG_SUM, 0, F64, F64
specs[1].fn, specs[1].input, specs[1].in_type, specs[1].out_type = capi.AGG_COUNT, -1, I64, capi.UINT64
kt, kn = (C.c_int32 * 1)(I64), (C.c_int32 * 1)(0)
in_cols = bench._cols(capi, [(p_, None, t) for p_, t in zip(d_in, types)])


def once():
    g = C.c_void_p()
    ctx.check(lib.ssb_group_create(ctx.h, 1, kt, kn, 2, specs, groups, C.byref(g)))
    ctx.check(lib.ssb_group_update_program(g, prog.h, in_cols, rows))
    ng = C.c_int64()
    ko, ao = (capi.Column * 1)(), (capi.Column * 2)()
    ctx.check(lib.ssb_group_finalize(g, C.byref(ng), ko, ao))
    cnt = np.zeros(ng.value, dtype=np.uint64)
    sums = np.zeros(ng.value, dtype=np.float64)
    ctx.d2h(cnt, ao[1].data)
    ctx.d2h(sums, ao[0].data)
    lib.ssb_group_destroy(g)
    return ng.value, int(cnt.sum()), float(sums.sum())


res = {}
for mode in ("0", "1"):
    os.environ["SSB200_GROUP_JIT"] = mode
    once()
    best = 1e9
    for _ in range(3):
        ctx.sync()
        t0 = time.perf_counter()
        r = once()
        ctx.sync()
        best = min(best, time.perf_counter() - t0)
    res[mode] = r
    print("jit=%s rows=%d groups=%d  %.3f ms  %.2f G rows/s  (groups %d, kept %d)" % (mode, rows, groups, best * 1e3, rows / best / 1e9, r[0], r[1]), flush=True)
assert res["0"][:2] == res["1"][:2] and abs(res["0"][2] - res["1"][2]) <= 1e-9 * abs(res["0"][2]), res
