"""The sharded C4 join alone (torchrun, N ranks), with a plain local join of the same size timed before and after it:
python -m torch.distributed.run --nproc-per-node N tools/micro_shard_join.py"""
import ctypes as C
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from supersonic_b200 import capi
from supersonic_b200.distributed import make_comm
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
ctx = capi.Context(int(os.environ["LOCAL_RANK"]))
lib = ctx.lib
P, B = 125_000_000, 12_500_000


def local_join(tag):
    pk, fk = ctx.malloc(B * 8 + 256), ctx.malloc(P * 8 + 256)
    ctx.generate(pk, B, 0, 7, 40, 4, 1000000007, B)
    ctx.generate(fk, P, 0, 7, 42, 1, 0, B)
    ctx.sync()
    for flags in (1, 1 | 0x100):
        for it in range(2):
            t0 = time.perf_counter()
            j = C.c_void_p()
            ctx.check(lib.ssb_join_build(ctx.h, 1, bench._cols(capi, [(pk, None, capi.INT64)]), B, flags, C.byref(j)))
            ctx.sync()
            t1 = time.perf_counter()
            n, pl, pr = C.c_int64(), C.c_void_p(), C.c_void_p()
            ctx.check(lib.ssb_join_probe(j, bench._cols(capi, [(fk, None, capi.INT64)]), P, 0, C.byref(n), C.byref(pl), C.byref(pr)))
            ctx.sync()
            t2 = time.perf_counter()
            lib.ssb_join_destroy(j)
        if rank == 0:
            print("[%s] local join, flags 0x%x: build %.3f ms, probe %.3f ms, pairs %d" % (tag, flags, (t1 - t0) * 1e3, (t2 - t1) * 1e3, n.value), flush=True)
    ctx.free(pk)
    ctx.free(fk)


local_join("before the communicator")
comm = make_comm(ctx)
local_join("after ssb_comm_create")
r = bench.hash_join_aux(capi, ctx, rank, world, P, B, dist, torch, comm)
if rank == 0:
    print(r["seconds"], r["exchange"])
local_join("after the sharded join")
comm.close()
dist.destroy_process_group()
