"""Narrow filter plans under different kernel variants (SSB200_EXPR_VARIANT): rows/s is the figure of merit."""
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from supersonic_b200 import capi
ctx = capi.Context(0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
n = capi.node; I64, BOOL = capi.INT64, capi.BOOL
cols = []
for i in range(4):
    p = ctx.malloc(rows*8+256); ctx.generate(p, rows, 0, 42, i, 0, [-(1<<31),-(1<<31),-(1<<62),0][i], [1<<32,1<<32,1<<63,1<<20][i]); cols.append(p)
out = [ctx.malloc(rows*8+256) for _ in range(4)]
dcount = ctx.malloc(8); ctx.sync()
def run(name, nodes, used, outs, pred, bpr):
    prog = capi.Program(ctx, nodes, [I64]*len(used), [0]*len(used), outs, predicate=pred)
    ins = [(cols[i], None, I64) for i in used]
    os_ = [(out[j], None, I64) for j in range(len(outs))]
    for _ in range(3): prog.run(ins, rows, os_, dcount)
    ctx.sync(); ctx.timer_start()
    for _ in range(5): prog.run(ins, rows, os_, dcount)
    ms = ctx.timer_stop()/5
    k=np.zeros(1,dtype=np.int64); ctx.d2h(k,dcount); sel=k[0]/rows
    print("variant=%s %-26s %.3f ms  %.1f Grows/s  %.1f GB/s" % (os.environ.get("SSB200_EXPR_VARIANT","default"), name, ms, rows/ms/1e6, bpr(sel)*rows/ms/1e6)); sys.stdout.flush()
    prog.close()
inp=lambda i: n(capi.OP_INPUT,I64,[i])
run("filter d<K -> d", [inp(0), n(capi.OP_CONST,I64,[],i64=1<<19), n(capi.OP_LT,BOOL,[0,1])], [3], [0], 2, lambda s: 8+8*s)
run("filter d<K -> a", [inp(0), inp(1), n(capi.OP_CONST,I64,[],i64=1<<19), n(capi.OP_LT,BOOL,[1,2])], [0,3], [0], 3, lambda s: 16+8*s)
nodes=[inp(0),inp(1),inp(2),inp(3),n(capi.OP_MUL,I64,[0,1]),n(capi.OP_ADD,I64,[4,2]),n(capi.OP_CONST,I64,[],i64=1<<19),n(capi.OP_LT,BOOL,[3,6])]
run("C2 filter+project", nodes, [0,1,2,3], [5], 7, lambda s: 32+8*s)
