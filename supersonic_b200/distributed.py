"""Row-range sharding across the GPUs of one box.

Compute / Project / Filter need no exchange: every rank runs the fused kernel on its own row
range and the ordered result is the concatenation in rank order. Aggregates need one exchange
step: every rank aggregates its shard into a dense partial table (ssb_group_update/finalize),
the partials are all-gathered (NCCL over NVLink when the tensors are CUDA tensors, gloo in the
CPU tests of this plumbing) and merged with ssb_group_merge (SUM adds, COUNT adds, MIN/MAX keep
the extreme, all-NULL partials stay NULL). One process per GPU; torch.distributed is plumbing.
"""
import ctypes as C

import numpy as np


def shard_rows(total_rows, rank, world, align=1024):
    """Contiguous row range [begin, end) of `rank`; interior boundaries are multiples of
    `align` rows so that null-bitmap words never straddle two shards."""
    def cut(r):
        if r >= world:
            return total_rows
        b = (total_rows * r) // world
        return min(total_rows, (b // align) * align)
    return cut(rank), cut(rank + 1)


def allgather_ragged(tensor, group=None):
    """All-gathers 1-D tensors of different lengths; returns the list in rank order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([tensor.numel()], dtype=torch.int64, device=tensor.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    padded = torch.zeros(cap, dtype=tensor.dtype, device=tensor.device)
    padded[:tensor.numel()] = tensor
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return [o[:s] for o, s in zip(out, sizes)]


class ShardedGroupAggregate(object):
    """GroupAggregate over a row-range sharded table: local GPU aggregation, all-gather of the
    dense partial tables, merge on every rank. Single-key, non-nullable INT64 keys (the C3
    shape); aggregates as (fn, in_type, out_type) over one value column each."""

    def __init__(self, ctx, aggs, expected_groups=0):
        from supersonic_b200 import capi
        self.capi, self.ctx, self.aggs, self.expected = capi, ctx, aggs, expected_groups

    def _create(self):
        capi, ctx = self.capi, self.ctx
        specs = (capi.AggSpec * len(self.aggs))()
        vi = 0
        for i, (fn, in_type, out_type, has_input) in enumerate(self.aggs):
            specs[i].fn, specs[i].in_type, specs[i].out_type = fn, in_type, out_type
            specs[i].input = vi if has_input else -1
            vi += 1 if has_input else 0
        kt, kn = (C.c_int32 * 1)(capi.INT64), (C.c_int32 * 1)(0)
        g = C.c_void_p()
        ctx.check(ctx.lib.ssb_group_create(ctx.h, 1, kt, kn, len(self.aggs), specs, self.expected, C.byref(g)))
        return g

    def _cols(self, items):
        arr = (self.capi.Column * max(1, len(items)))()
        for i, (d, n, t) in enumerate(items):
            arr[i].data, arr[i].nulls, arr[i].dtype = d, n, t
        return arr

    def local(self, key_ptr, value_cols, rows):
        """Aggregates this rank's shard. value_cols: [(ptr, nulls_ptr, dtype)]. Returns the group
        handle and (n_groups, key column, agg columns) living in device memory owned by it."""
        capi, ctx = self.capi, self.ctx
        g = self._create()
        ctx.check(ctx.lib.ssb_group_update(g, self._cols([(key_ptr, None, capi.INT64)]), self._cols(value_cols), rows))
        n = C.c_int64()
        ko = self._cols([(0, None, 0)])
        ao = self._cols([(0, None, 0)] * len(self.aggs))
        ctx.check(ctx.lib.ssb_group_finalize(g, C.byref(n), ko, ao))
        return g, n.value, ko, ao

    def merge_into(self, g, n_groups, key_ptr, agg_ptrs):
        """Adds a (gathered) dense partial table into the group handle `g` (ssb_group_merge)."""
        capi, ctx = self.capi, self.ctx
        keys = self._cols([(key_ptr, None, capi.INT64)])
        aggs = self._cols([(p, None, self.aggs[i][2]) for i, p in enumerate(agg_ptrs)])
        ctx.check(ctx.lib.ssb_group_merge(g, n_groups, keys, aggs))

    def run(self, key_ptr, value_cols, rows, group=None):
        """Single-rank aggregate. Returns (keys, [agg arrays], [agg null masks]) as numpy."""
        capi, ctx = self.capi, self.ctx
        g, n, ko, ao = self.local(key_ptr, value_cols, rows)
        # single rank: the local table is the result
        keys = np.zeros(n, dtype=np.int64)
        if n:
            ctx.d2h(keys, ko[0].data)
        outs, nulls = [], []
        for i, (fn, in_type, out_type, _) in enumerate(self.aggs):
            a = np.zeros(n, dtype=np.float64 if out_type == capi.DOUBLE else (np.uint64 if out_type == capi.UINT64 else np.int64))
            if n:
                ctx.d2h(a, ao[i].data)
            outs.append(a)
            nulls.append(None)
        ctx.lib.ssb_group_destroy(g)
        return keys, outs, nulls
