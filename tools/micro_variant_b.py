"""BASELINE config 2 variant B (ProjectAllAttributes: eight columns + e = a*b+c under the filter) timed with a
stopwatch around the whole program (a wide plan runs as several column-group kernels)."""
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from supersonic_b200 import capi
ctx = capi.Context(0)
rows = 100_000_000
n = capi.node; I64, BOOL = capi.INT64, capi.BOOL
cols=[]
for i in range(8):
    p = ctx.malloc(rows*8+256); ctx.generate(p, rows, 0, 42, i, 0, [-(1<<31),-(1<<31),-(1<<62),0,0,0,0,0][i], [1<<32,1<<32,1<<63,1<<20,0,0,0,0][i]); cols.append(p)
out = [ctx.malloc(rows*8+256) for _ in range(9)]
dcount = ctx.malloc(8); ctx.sync(); ctx.enable_timing(True)
inp=lambda i: n(capi.OP_INPUT,I64,[i])
nodes=[inp(i) for i in range(8)]+[n(capi.OP_MUL,I64,[0,1]),n(capi.OP_ADD,I64,[8,2]),n(capi.OP_CONST,I64,[],i64=1<<19),n(capi.OP_LT,BOOL,[3,10])]
prog = capi.Program(ctx, nodes, [I64]*8, [0]*8, [9,0,1,2,3,4,5,6,7], predicate=11)
ins=[(cols[i],None,I64) for i in range(8)]; os_=[(out[j],None,I64) for j in range(9)]
for _ in range(3): prog.run(ins, rows, os_, dcount)
ctx.sync(); ctx.timer_start()
for _ in range(5): prog.run(ins, rows, os_, dcount)
ms=ctx.timer_stop()/5
print("variant B %d rows: %.3f ms  %.1f GB/s" % (rows, ms, rows*100/ms/1e6))
