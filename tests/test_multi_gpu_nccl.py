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


@pytest.mark.parametrize("strategy", ["all_to_all", "broadcast"])
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


@pytest.mark.parametrize("strategy", ["all_to_all", "broadcast"])
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
