"""One-GPU check of ShardedSort through the real kernels (world size 1 over NCCL): every case against numpy."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
import torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
from supersonic_b200 import capi
from supersonic_b200.distributed import CudaJoinKernels, ShardedSort
from test_multi_gpu import _sort_table, SORT_CASES, NumpyJoinKernels
kern = CudaJoinKernels(capi.Context(0))
t = _sort_table()
for keys, desc in SORT_CASES:
    col = lambda nm, dt: (torch.from_numpy(np.ascontiguousarray(t[nm])).cuda(), dt)
    ks, cs = ShardedSort(kern).run([col(nm, dt) for nm, dt in keys], desc, [col("id", 2), col("v", 2)])
    want = NumpyJoinKernels().sort_perm([(torch.from_numpy(t[nm]), dt) for nm, dt in keys], desc).numpy()
    assert np.array_equal(cs[0][0].cpu().numpy(), t["id"][want]), (keys, desc)
    assert np.array_equal(cs[1][0].cpu().numpy(), t["v"][want])
    for (k, _), (nm, _) in zip(ks, keys):
        assert np.array_equal(k.cpu().numpy(), t[nm][want])
print("sharded sort (1 GPU, nccl): %d cases OK, launches=%d" % (len(SORT_CASES), kern.ctx.launches()))
dist.destroy_process_group()
