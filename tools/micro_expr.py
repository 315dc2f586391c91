import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from supersonic_b200 import capi
ctx = capi.Context(0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
n = capi.node; I64, BOOL = capi.INT64, capi.BOOL
cols = []
for i in range(8):
    p = ctx.malloc(rows*8+256); ctx.generate(p, rows, 0, 42, i, 0, [-(1<<31),-(1<<31),-(1<<62),0,0,0,0,0][i], [1<<32,1<<32,1<<63,1<<20,0,0,0,0][i]); cols.append(p)
out = [ctx.malloc(rows*8+256) for _ in range(9)]
dcount = ctx.malloc(8)
ctx.sync(); ctx.enable_timing(True)
def run(name, nodes, nin, outs, pred, bytes_per_row):
    prog = capi.Program(ctx, nodes, [I64]*nin, [0]*nin, outs, predicate=pred)
    ins = [(cols[i], None, I64) for i in range(nin)]
    os_ = [(out[j], None, I64) for j in range(len(outs))]
    for _ in range(3): prog.run(ins, rows, os_, dcount)
    ms=[]
    for _ in range(5):
        prog.run(ins, rows, os_, dcount); ms.append(ctx.last_kernel_ms())
    k=np.zeros(1,dtype=np.int64); ctx.d2h(k,dcount)
    m=float(np.median(ms)); sel=k[0]/rows
    b=bytes_per_row(sel)*rows
    print("%-28s %.3f ms  %.1f GB/s  kept=%.3f" % (name, m, b/m/1e6, sel)); sys.stdout.flush()
    prog.close()
inp=lambda i: n(capi.OP_INPUT,I64,[i])
# 1 copy one column
run("copy a", [inp(0)], 1, [0], -1, lambda s: 16)
run("copy a,b,c,d", [inp(0),inp(1),inp(2),inp(3)], 4, [0,1,2,3], -1, lambda s: 64)
run("compute a*b+c", [inp(0),inp(1),inp(2),n(capi.OP_MUL,I64,[0,1]),n(capi.OP_ADD,I64,[3,2])], 3, [4], -1, lambda s: 32)
run("filter d<K -> d", [inp(0), n(capi.OP_CONST,I64,[],i64=1<<19), n(capi.OP_LT,BOOL,[0,1])], 1, [0], 2, lambda s: 8+8*s) if False else None
# predicate on column d (index 3)
nodes=[inp(0),inp(1),inp(2),inp(3),n(capi.OP_MUL,I64,[0,1]),n(capi.OP_ADD,I64,[4,2]),n(capi.OP_CONST,I64,[],i64=1<<19),n(capi.OP_LT,BOOL,[3,6])]
run("C2 filter+project", nodes, 4, [5], 7, lambda s: 32+8*s)
nodes=[inp(0),inp(1),inp(2),inp(3),n(capi.OP_MUL,I64,[0,1]),n(capi.OP_ADD,I64,[4,2]),n(capi.OP_CONST,I64,[],i64=1<<10),n(capi.OP_LT,BOOL,[3,6])]
run("C2 sel=0.001", nodes, 4, [5], 7, lambda s: 32+8*s)
nodes=[inp(0),inp(1),inp(2),inp(3),n(capi.OP_CONST,I64,[],i64=1<<19),n(capi.OP_LT,BOOL,[3,4])]
run("filter d<K -> a (2 cols)", nodes, 4, [0], 5, lambda s: 16+8*s)

# selectivity sweep of the C2 plan (s in {0.01, 0.5, 0.99})
for label, k in [("C2 sel=0.01", 10486), ("C2 sel=0.50", 1 << 19), ("C2 sel=0.99", 1038090)]:
    nodes=[inp(0),inp(1),inp(2),inp(3),n(capi.OP_MUL,I64,[0,1]),n(capi.OP_ADD,I64,[4,2]),n(capi.OP_CONST,I64,[],i64=k),n(capi.OP_LT,BOOL,[3,6])]
    run(label, nodes, 4, [5], 7, lambda s: 32+8*s)
# variant B: ProjectAllAttributes (the 8 columns a..h) + e = a*b+c under the same filter: 64 B in, 72 B out per kept row
nodes=[inp(i) for i in range(8)]+[n(capi.OP_MUL,I64,[0,1]),n(capi.OP_ADD,I64,[8,2]),n(capi.OP_CONST,I64,[],i64=1<<19),n(capi.OP_LT,BOOL,[3,10])]
run("C2 variant B (9 outputs)", nodes, 8, [9,0,1,2,3,4,5,6,7], 11, lambda s: 64+72*s)
