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
# nullable keys and payload
from test_multi_gpu import _null_sort_table, NULL_SORT_CASES
t = _null_sort_table()
dev = lambda nm: torch.from_numpy(np.ascontiguousarray(t[nm])).cuda()
cpu = lambda nm: torch.from_numpy(np.ascontiguousarray(t[nm]))
for desc in NULL_SORT_CASES:
    ks, cs, kn, cn = ShardedSort(kern).run([(dev("k"), 2), (dev("j"), 2)], desc, [(dev("id"), 2), (dev("w"), 5)],
                                           key_nulls=[dev("k_null"), dev("j_null")], col_nulls=[None, dev("w_null")])
    want = NumpyJoinKernels().sort_perm([(cpu("k"), 2), (cpu("j"), 2)], desc, [cpu("k_null"), cpu("j_null")]).numpy()
    assert np.array_equal(cs[0][0].cpu().numpy(), t["id"][want]), desc
    assert np.array_equal(kn[0].cpu().numpy(), t["k_null"][want]) and np.array_equal(cn[1].cpu().numpy(), t["w_null"][want])
print("sharded sort (1 GPU, nccl): %d cases OK, launches=%d" % (len(SORT_CASES), kern.ctx.launches()))
dist.destroy_process_group()
