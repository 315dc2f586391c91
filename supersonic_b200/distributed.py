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


def allgather_tables(table, group=None):
    """All-gathers the per-rank hash tables (1-D int64 tensors, usually of different lengths) without padding
    copies of the big tensors: sizes first, then one broadcast per rank straight into its final tensor."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([table.numel()], dtype=torch.int64, device=table.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    if len(set(sizes)) == 1 and hasattr(dist, "all_gather_into_tensor") and table.is_cuda:
        out = torch.empty(world * sizes[0], dtype=table.dtype, device=table.device)
        dist.all_gather_into_tensor(out, table.contiguous(), group=group)
        return [out[r * sizes[0]:(r + 1) * sizes[0]] for r in range(world)]
    out = []
    for r in range(world):
        t = table.contiguous() if r == rank else torch.empty(sizes[r], dtype=table.dtype, device=table.device)
        src = dist.get_global_rank(group, r) if group is not None else r
        dist.broadcast(t, src=src, group=group)
        out.append(t)
    return out


class _DevArray(object):
    """Zero-copy view of device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


_TYPESTR = {1: "<i4", 2: "<i8", 3: "<i8", 4: "<i8", 5: "<f8", 6: "|u1", 8: "<i4", 9: "<f4", 10: "<i4", 13: "<i4"}


def make_comm(ctx, group=None):
    """A libssb200 communicator (ssb_comm_*, NCCL inside the library, on the context's stream) for the
    ranks of a torch.distributed process group: torch only carries the 128-byte id from rank 0."""
    import torch
    import torch.distributed as dist
    from supersonic_b200 import capi
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if dist.get_backend(group) == "nccl":
        ident = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident.copy_(torch.frombuffer(bytearray(capi.Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(ident.cpu().numpy().tobytes())
    else:
        box = [capi.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        raw = box[0]
    return capi.Comm(ctx, raw, world, rank)


def merge_group_partials(ctx, g, key_types, agg_types, group=None, comm=None):
    """The exchange step of a row-range sharded GroupAggregate / ScalarAggregate:
    every rank finalizes its table `g` into dense partial columns, the partials are all-gathered
    (NCCL) and the other ranks' rows are added with ssb_group_merge; after the call every rank
    holds the whole-table result. NOT NULL keys and aggregates. Returns (n_groups, key columns,
    aggregate columns) as ctypes Column arrays owned by `g`."""
    import torch
    import torch.distributed as dist
    from supersonic_b200 import capi
    lib = ctx.lib

    def cols(n):
        return (capi.Column * max(1, n))()

    n = C.c_int64()
    ko, ao = cols(len(key_types)), cols(len(agg_types))
    if comm is not None:
        # the library's own exchange: reduce-scatter by key hash + all-gather (ssb_shard_group_merge); NULL
        # keys and all-NULL aggregates travel with their is_null flags
        ctx.check(lib.ssb_shard_group_merge(comm.h, g, C.byref(n), ko, ao))
        return n.value, ko, ao
    ctx.check(lib.ssb_group_finalize(g, C.byref(n), ko, ao))
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return n.value, ko, ao
    rank = dist.get_rank(group)

    def view(col, dtype):
        return torch.as_tensor(_DevArray(col.data, n.value, _TYPESTR[dtype]), device="cuda") if n.value else \
            torch.empty(0, dtype=_torch_dtype(dtype), device="cuda")

    # this torch-collective form carries no is_null flags: refuse tables whose partials can be NULL instead of
    # merging placeholder values; ssb_shard_group_merge (comm=...) handles them
    if any(ko[i].nulls for i in range(len(key_types))):
        raise ValueError("merge_group_partials without a communicator needs NOT NULL keys; pass comm=make_comm(ctx)")
    gk = [allgather_ragged(view(ko[i], t), group) for i, t in enumerate(key_types)]
    ga = [allgather_ragged(view(ao[i], t), group) for i, t in enumerate(agg_types)]
    torch.cuda.synchronize()
    for r in range(world):
        if r == rank:
            continue
        rows = (gk[0][r] if key_types else ga[0][r]).numel()
        if rows == 0:
            continue
        kc, ac = cols(len(key_types)), cols(len(agg_types))
        for i, t in enumerate(key_types):
            kc[i].data, kc[i].nulls, kc[i].dtype = gk[i][r].data_ptr(), None, t
        for i, t in enumerate(agg_types):
            ac[i].data, ac[i].nulls, ac[i].dtype = ga[i][r].data_ptr(), None, t
        ctx.check(lib.ssb_group_merge(g, rows, kc, ac))
    ctx.check(lib.ssb_group_finalize(g, C.byref(n), ko, ao))
    return n.value, ko, ao


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


# ---------------------------------------------------------------------------------------------
# Hash join over row-range shards (BASELINE config 4)
# ---------------------------------------------------------------------------------------------
INNER, LEFT_OUTER = 0, 1
NOT_UNIQUE, UNIQUE = 0, 1


def _torch_dtype(dtype):
    """torch container of an SSB_* column type (unsigned columns travel as their signed bit image)."""
    import torch
    from supersonic_b200 import capi
    return {capi.INT64: torch.int64, capi.DATETIME: torch.int64, capi.UINT64: torch.int64,
            capi.DOUBLE: torch.float64, capi.INT32: torch.int32, capi.DATE: torch.int32,
            capi.ENUM: torch.int32, capi.UINT32: torch.int32, capi.FLOAT: torch.float32,
            capi.BOOL: torch.uint8}[dtype]


class _JoinHandle(object):
    """Keeps an ssb_join alive while a tensor views its table."""

    def __init__(self, lib, h):
        self.lib, self.h = lib, h

    def __del__(self):
        if self.h:
            self.lib.ssb_join_destroy(self.h)
            self.h = None


class CudaJoinKernels(object):
    """The four data-path steps of the sharded join, each one C-ABI call into libssb200.so on
    the device pointers of torch tensors (torch: allocation and collectives only). The torch
    work of the join is issued on the context's own stream (`stream`), so kernels and
    collectives stay ordered without host synchronisation."""

    def __init__(self, ctx):
        import torch
        from supersonic_b200 import capi
        self.torch, self.capi, self.ctx = torch, capi, ctx
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(ctx.lib.ssb_ctx_stream(ctx.h), device=self.device)

    def _cols(self, cols):
        arr = (self.capi.Column * max(1, len(cols)))()
        for i, (t, dtype) in enumerate(cols):
            arr[i].data, arr[i].nulls, arr[i].dtype = t.data_ptr(), None, dtype
        return arr

    def empty(self, n, dtype):
        return self.torch.empty(max(int(n), 0), dtype=_torch_dtype(dtype), device=self.device)

    def partition(self, keys, n_parts, null_part):
        """Stable hash partition: (row ids grouped by part, rows per part)."""
        rows = keys[0][0].numel()
        perm = self.empty(rows, self.capi.INT64)
        counts = (C.c_int64 * n_parts)()
        self.ctx.check(self.ctx.lib.ssb_partition_rows(self.ctx.h, len(keys), self._cols(keys), rows, n_parts,
                                                       null_part, perm.data_ptr(), counts))
        return perm, [int(c) for c in counts]

    def gather(self, col, idx, want_valid=False):
        """out[i] = col[idx[i]]; with want_valid also an is_null byte per row: 1 where idx[i] < 0."""
        t, dtype = col
        n = idx.numel()
        out = self.empty(n, dtype)
        dst = self._cols([(out, dtype)])
        bitmap = None
        if want_valid:
            bitmap = self.torch.zeros((n + 31) // 32 + 1, dtype=self.torch.int32, device=self.device)
            dst[0].nulls = bitmap.data_ptr()
        if n:
            self.ctx.check(self.ctx.lib.ssb_gather(self.ctx.h, self._cols([col]), idx.data_ptr(), n, dst))
        if not want_valid:
            return out
        is_null = self.torch.zeros(n, dtype=self.torch.uint8, device=self.device)
        if n:
            self.ctx.check(self.ctx.lib.ssb_nulls_unpack(self.ctx.h, bitmap.data_ptr(), n, is_null.data_ptr()))
        return out, is_null

    def join(self, build_keys, probe_keys, join_type, uniqueness):
        """(lhs row, rhs row) pairs of the local join in lhs order (rhs -1: unmatched, LEFT_OUTER)."""
        lib, ctx, torch = self.ctx.lib, self.ctx, self.torch
        h = C.c_void_p()
        ctx.check(lib.ssb_join_build(ctx.h, len(build_keys), self._cols(build_keys), build_keys[0][0].numel(),
                                     uniqueness, C.byref(h)))
        try:
            n = C.c_int64()
            pl, pr = C.c_void_p(), C.c_void_p()
            ctx.check(lib.ssb_join_probe(h, self._cols(probe_keys), probe_keys[0][0].numel(), join_type,
                                         C.byref(n), C.byref(pl), C.byref(pr)))
            li, ri = self.empty(n.value, self.capi.INT64), self.empty(n.value, self.capi.INT64)
            if n.value:
                ctx.check(lib.ssb_memcpy_d2d(ctx.h, li.data_ptr(), pl, n.value * 8))
                ctx.check(lib.ssb_memcpy_d2d(ctx.h, ri.data_ptr(), pr, n.value * 8))
                ctx.sync()   # the pair buffers die with the join handle
        finally:
            lib.ssb_join_destroy(h)
        return li, ri

    def build_table(self, build_keys):
        """Index over UNIQUE single-column keys as a tensor that can travel: the table's slots (16 bytes each:
        key word, build row + 1) viewed as int64. Returns (table tensor, keep-alive handle)."""
        lib, ctx, torch = self.ctx.lib, self.ctx, self.torch
        h = C.c_void_p()
        ctx.check(lib.ssb_join_build(ctx.h, len(build_keys), self._cols(build_keys), build_keys[0][0].numel(),
                                     UNIQUE | 0x100, C.byref(h)))   # SSB_KEYS_COMPACT_TABLE: the table travels
        slots, cap = C.c_void_p(), C.c_int64()
        ctx.check(lib.ssb_join_table(h, C.byref(slots), C.byref(cap)))
        view = torch.as_tensor(_DevArray(slots.value, 2 * cap.value, "<i8"), device=self.device)
        return view, _JoinHandle(lib, h)

    def probe_parts(self, tables, row_offsets, probe_keys, join_type):
        """Probes the gathered per-part tables (ssb_join_attach_parts): (lhs row, row_offsets[part] + rhs row)
        pairs in lhs order; rhs -1 for unmatched rows of a LEFT_OUTER join."""
        lib, ctx = self.ctx.lib, self.ctx
        n_parts = len(tables)
        ptrs = (C.c_void_p * n_parts)(*[t.data_ptr() for t in tables])
        caps = (C.c_int64 * n_parts)(*[t.numel() // 2 for t in tables])
        offs = (C.c_int64 * n_parts)(*[int(o) for o in row_offsets])
        h = C.c_void_p()
        ctx.check(lib.ssb_join_attach_parts(ctx.h, probe_keys[0][1], n_parts, ptrs, caps, offs, C.byref(h)))
        try:
            n = C.c_int64()
            pl, pr = C.c_void_p(), C.c_void_p()
            ctx.check(lib.ssb_join_probe(h, self._cols(probe_keys), probe_keys[0][0].numel(), join_type,
                                         C.byref(n), C.byref(pl), C.byref(pr)))
            li, ri = self.empty(n.value, self.capi.INT64), self.empty(n.value, self.capi.INT64)
            if n.value:
                ctx.check(lib.ssb_memcpy_d2d(ctx.h, li.data_ptr(), pl, n.value * 8))
                ctx.check(lib.ssb_memcpy_d2d(ctx.h, ri.data_ptr(), pr, n.value * 8))
                ctx.sync()
        finally:
            lib.ssb_join_destroy(h)
        return li, ri

    def scatter(self, col, idx, dst):
        """dst[idx[i]] = col[i] (distinct indices)."""
        t, dtype = col
        if idx.numel():
            self.ctx.check(self.ctx.lib.ssb_scatter(self.ctx.h, self._cols([col]), idx.data_ptr(), idx.numel(),
                                                    self._cols([(dst, dtype)])))

    def zeros(self, n, dtype):
        return self.torch.zeros(max(int(n), 0), dtype=_torch_dtype(dtype), device=self.device)

    def iota(self, n):
        return self.torch.arange(max(int(n), 0), dtype=self.torch.int64, device=self.device)

    def compact(self, flag, cols):
        """Rows of `cols` where flag != 0, in order: the fused Filter kernel with the flag column as
        predicate (ssb_program_run), at most ten columns per launch."""
        capi = self.capi
        n = flag.numel()
        out, kept = [], 0
        for first in range(0, len(cols), 10):
            part = cols[first:first + 10]
            nodes = [capi.node(capi.OP_INPUT, capi.BOOL, [0])]
            nodes += [capi.node(capi.OP_INPUT, dt, [i + 1]) for i, (_, dt) in enumerate(part)]
            prog = capi.Program(self.ctx, nodes, [capi.BOOL] + [dt for _, dt in part], [0] * (len(part) + 1),
                                list(range(1, len(part) + 1)), predicate=0)
            outs = [self.empty(n, dt) for _, dt in part]
            kept = prog.run_sync([(flag.data_ptr(), None, capi.BOOL)] + [(t.data_ptr(), None, dt) for t, dt in part], n,
                                 [(o.data_ptr(), None, dt) for o, (_, dt) in zip(outs, part)]) if n else 0
            prog.close()
            out += [(o[:kept], dt) for o, (_, dt) in zip(outs, part)]
        return out

    def scope(self):
        """Context manager: torch work issued inside runs on the library's stream."""
        return self.torch.cuda.stream(self.stream)

    def finish(self):
        self.ctx.sync()

    def sort_perm(self, keys, descending, key_nulls=None):
        """Stable sort permutation by several key columns (first = most significant); key_nulls: per key
        a uint8 tensor (1 = NULL) or None -- NULLs first for ASCENDING, last for DESCENDING."""
        n = keys[0][0].numel()
        perm = self.empty(n, self.capi.INT64)
        if n:
            desc = (C.c_int32 * len(keys))(*[1 if d else 0 for d in descending])
            cols = self._cols(keys)
            bitmaps = []    # keep the packed bitmaps alive until the sort has run
            for i, flags in enumerate(key_nulls or []):
                if flags is None:
                    continue
                words = self.torch.zeros(n // 32 + 2, dtype=self.torch.int32, device=self.device)
                self.ctx.check(self.ctx.lib.ssb_nulls_pack(self.ctx.h, flags.data_ptr(), n, words.data_ptr()))
                cols[i].nulls = words.data_ptr()
                bitmaps.append(words)
            self.ctx.check(self.ctx.lib.ssb_sort_permutation(self.ctx.h, len(keys), cols, desc, n, perm.data_ptr()))
        return perm

    def order_by(self, key):
        """Stable ascending order of an INT64 column."""
        n = key.numel()
        perm = self.empty(n, self.capi.INT64)
        if n:
            desc = (C.c_int32 * 1)(0)
            self.ctx.check(self.ctx.lib.ssb_sort_permutation(self.ctx.h, 1, self._cols([(key, self.capi.INT64)]), desc,
                                                             n, perm.data_ptr()))
        return perm


def _exchange(tensor, send_counts, recv_counts, group=None):
    """all-to-all of one column: rank r receives, in rank order, the slice every rank cut for r."""
    import torch
    import torch.distributed as dist
    out = torch.empty(int(sum(recv_counts)), dtype=tensor.dtype, device=tensor.device)
    dist.all_to_all_single(out, tensor.contiguous(), [int(c) for c in recv_counts], [int(c) for c in send_counts],
                           group=group)
    return out


def _exchange_counts(send_counts, device, group=None):
    import torch
    import torch.distributed as dist
    s = torch.tensor(send_counts, dtype=torch.int64, device=device)
    r = torch.empty_like(s)
    dist.all_to_all_single(r, s, group=group)
    return [int(x) for x in r.tolist()]


class ShardedHashJoin(object):
    """HashJoin(INNER | LEFT_OUTER, UNIQUE | NOT_UNIQUE) over tables sharded by row range.

    Both sides are redistributed by key hash with one all-to-all each (NCCL over NVLink for CUDA
    tensors): the build side with its payload columns, the probe side as keys + origin row id
    only. Every rank joins what it received (ssb_join_build / ssb_join_probe), gathers the rhs
    payload of its pairs and returns them to the rank owning the lhs row with a third
    all-to-all; the owner restores lhs order with a stable sort on the origin row id. A build
    row list received in rank order is in global insertion order and all pairs of one lhs row
    come from one rank, so the concatenation of the per-rank results in rank order is exactly
    the reference's output order (cursor/core/hash_join.cc:793-831).
    NULL keys and nullable payload columns: see run().

    Strategy "broadcast" instead all-gathers the build side (keys + payload, rank order = global
    insertion order), builds the whole table on every rank and probes the local lhs shard: no
    probe row moves and no order has to be restored. "auto" takes it when the whole build side
    stays under `broadcast_max_rows` rows (8x less traffic for 1B joined to 100M)."""

    def __init__(self, kernels, group=None, strategy="all_to_all", broadcast_max_rows=1 << 28):
        self.k, self.group, self.strategy, self.broadcast_max_rows = kernels, group, strategy, broadcast_max_rows

    def run(self, lhs_keys, lhs_cols, rhs_keys, rhs_cols, join_type=INNER, uniqueness=UNIQUE,
            lhs_key_nulls=None, rhs_key_nulls=None):
        """Columns are (tensor, SSB dtype) pairs of this rank's shards. Returns
        (lhs_rows, lhs payload columns, rhs payload columns, rhs_is_null or None): the join
        result for this rank's lhs shard in lhs order; lhs_rows are shard-local row ids.

        NULLs. `lhs_key_nulls` / `rhs_key_nulls`: one byte per row, 1 = some key column of the row is
        NULL. Such rows never match (hash_join.cc:67-76,616-617,755-756): build rows are dropped
        before the exchange, probe rows stay at home and come out unmatched (LEFT_OUTER) or not at
        all (INNER). A nullable payload column travels as two columns: its values and its is_null
        bytes as an extra BOOL column; for LEFT_OUTER the caller ORs those with rhs_is_null."""
        import torch
        k = self.k
        I64 = 2
        if rhs_key_nulls is not None:
            with k.scope():
                kept = k.compact(1 - rhs_key_nulls, list(rhs_keys) + list(rhs_cols))
            rhs_keys, rhs_cols = kept[:len(rhs_keys)], kept[len(rhs_keys):]
        if lhs_key_nulls is None:
            return self._run_not_null(lhs_keys, lhs_cols, rhs_keys, rhs_cols, join_type, uniqueness)
        n_l = lhs_keys[0][0].numel()
        with k.scope():
            kept = k.compact(1 - lhs_key_nulls, [(k.iota(n_l), I64)] + list(lhs_keys))
        sel, keys_nn = kept[0], kept[1:]
        rows_nn, _, out_r, r_null = self._run_not_null(keys_nn, [], rhs_keys, rhs_cols, join_type, uniqueness)
        with k.scope():
            rows = k.gather(sel, rows_nn)
            if join_type == LEFT_OUTER:
                # the rows with a NULL key come out once, unmatched, at their place in lhs order
                null_rows = k.compact(lhs_key_nulls, [(k.iota(n_l), I64)])[0][0]
                m = null_rows.numel()
                rows = torch.cat([rows, null_rows])
                order = k.order_by(rows)
                rows = k.gather((rows, I64), order)
                out_r = [(k.gather((torch.cat([t, k.zeros(m, dt)]), dt), order), dt) for t, dt in out_r]
                r_null = k.gather((torch.cat([r_null, k.zeros(m, 6) + 1]), 6), order)
            out_l = [(k.gather(c, rows), c[1]) for c in lhs_cols]
        k.finish()
        return rows, out_l, out_r, r_null

    def _run_not_null(self, lhs_keys, lhs_cols, rhs_keys, rhs_cols, join_type, uniqueness):
        import torch
        import torch.distributed as dist
        strategy = self.strategy
        if strategy == "auto":
            n = torch.tensor([rhs_keys[0][0].numel()], dtype=torch.int64, device=rhs_keys[0][0].device)
            dist.all_reduce(n, group=self.group)
            small = int(n.item()) <= self.broadcast_max_rows
            # UNIQUE single-column keys: every rank builds the table of one hash part only (1 / world of the
            # build work) and the tables are all-gathered; otherwise the whole build side is all-gathered
            strategy = ("replicate" if uniqueness == UNIQUE and len(rhs_keys) == 1 else "broadcast") if small else "all_to_all"
        with self.k.scope():
            if strategy == "replicate" and (uniqueness != UNIQUE or len(rhs_keys) != 1):
                strategy = "all_to_all"      # the replicated tables hold UNIQUE single-column keys
            if strategy == "broadcast":
                out = self._run_broadcast(lhs_keys, lhs_cols, rhs_keys, rhs_cols, join_type, uniqueness)
            elif strategy == "replicate":
                out = self._run_replicate(lhs_keys, lhs_cols, rhs_keys, rhs_cols, join_type)
            else:
                out = self._run(lhs_keys, lhs_cols, rhs_keys, rhs_cols, join_type, uniqueness)
        self.k.finish()
        self.last_strategy = strategy
        return out

    def _run_broadcast(self, lhs_keys, lhs_cols, rhs_keys, rhs_cols, join_type, uniqueness):
        import torch
        k, g = self.k, self.group
        gather = lambda c: (torch.cat(allgather_ragged(c[0], g)), c[1])   # noqa: E731
        b_keys = [gather(c) for c in rhs_keys]
        b_cols = [gather(c) for c in rhs_cols]
        li, ri = k.join(b_keys, lhs_keys, join_type, uniqueness)
        outer = join_type == LEFT_OUTER
        out_r, r_null = [], None
        for j, c in enumerate(b_cols):
            if outer and j == 0:
                vals, r_null = k.gather(c, ri, want_valid=True)
            else:
                vals = k.gather(c, ri)
            out_r.append((vals, c[1]))
        if outer and not b_cols:
            _, r_null = k.gather(b_keys[0], ri, want_valid=True)
        out_l = [(k.gather(c, li), c[1]) for c in lhs_cols]
        return li, out_l, out_r, r_null

    def _run_replicate(self, lhs_keys, lhs_cols, rhs_keys, rhs_cols, join_type):
        """UNIQUE single-column keys. The build side is redistributed by key hash (small: it is the build side),
        every rank builds the index of the part it received, the tables and the received payload columns are
        all-gathered, and the local lhs shard probes the table of each key's part: no probe row moves, the
        output is in lhs order, and the build work per rank is 1 / world of the whole (the broadcast form
        builds the whole table on every rank)."""
        import torch
        import torch.distributed as dist
        k, g = self.k, self.group
        world, rank = dist.get_world_size(g), dist.get_rank(g)
        device = lhs_keys[0][0].device
        perm_b, cnt_b = k.partition(rhs_keys, world, rank)
        rcv_b = _exchange_counts(cnt_b, device, g)
        b_keys = [(_exchange(k.gather(c, perm_b), cnt_b, rcv_b, g), c[1]) for c in rhs_keys]
        b_cols = [(_exchange(k.gather(c, perm_b), cnt_b, rcv_b, g), c[1]) for c in rhs_cols]
        table, keep = k.build_table(b_keys)
        tables = allgather_tables(table, g)
        parts_cols = [allgather_ragged(c[0], g) for c in b_cols]
        rows_per_part = [int(t.numel()) for t in allgather_ragged(b_keys[0][0], g)] if not parts_cols else \
            [int(t.numel()) for t in parts_cols[0]]
        offsets = [0]
        for r in rows_per_part[:-1]:
            offsets.append(offsets[-1] + r)
        all_cols = [(torch.cat(p), c[1]) for p, c in zip(parts_cols, b_cols)]
        li, ri = k.probe_parts(tables, offsets, lhs_keys, join_type)
        del keep
        outer = join_type == LEFT_OUTER
        out_r, r_null = [], None
        for j, c in enumerate(all_cols):
            if outer and j == 0:
                vals, r_null = k.gather(c, ri, want_valid=True)
            else:
                vals = k.gather(c, ri)
            out_r.append((vals, c[1]))
        if outer and not all_cols:
            _, r_null = k.gather((ri, 2), ri, want_valid=True)
        out_l = [(k.gather(c, li), c[1]) for c in lhs_cols]
        return li, out_l, out_r, r_null

    def _run(self, lhs_keys, lhs_cols, rhs_keys, rhs_cols, join_type, uniqueness):
        import torch.distributed as dist
        k, g = self.k, self.group
        world, rank = dist.get_world_size(g), dist.get_rank(g)
        device = lhs_keys[0][0].device
        I64 = 2
        # ---- build side: keys + payload to the rank that owns the key's hash range
        perm_b, cnt_b = k.partition(rhs_keys, world, rank)
        rcv_b = _exchange_counts(cnt_b, device, g)
        b_keys = [(_exchange(k.gather(c, perm_b), cnt_b, rcv_b, g), c[1]) for c in rhs_keys]
        b_cols = [(_exchange(k.gather(c, perm_b), cnt_b, rcv_b, g), c[1]) for c in rhs_cols]
        # ---- probe side: keys + origin row id
        perm_p, cnt_p = k.partition(lhs_keys, world, rank)
        rcv_p = _exchange_counts(cnt_p, device, g)
        p_keys = [(_exchange(k.gather(c, perm_p), cnt_p, rcv_p, g), c[1]) for c in lhs_keys]
        p_origin = _exchange(perm_p, cnt_p, rcv_p, g)
        # ---- local join; pairs come in received-lhs order = (source rank, origin row)
        li, ri = k.join(b_keys, p_keys, join_type, uniqueness)
        bounds = [0]
        for c in rcv_p:
            bounds.append(bounds[-1] + c)
        import torch
        edges = torch.searchsorted(li, torch.tensor(bounds, dtype=torch.int64, device=device)).tolist() \
            if li.numel() else [0] * (world + 1)
        back = [int(edges[s + 1] - edges[s]) for s in range(world)]
        rcv_back = _exchange_counts(back, device, g)
        origin = _exchange(k.gather((p_origin, I64), li), back, rcv_back, g)
        outer = join_type == LEFT_OUTER
        r_cols, r_null = [], None
        for j, c in enumerate(b_cols):
            if outer and j == 0:
                vals, is_null = k.gather(c, ri, want_valid=True)
                r_null = _exchange(is_null, back, rcv_back, g)
            else:
                vals = k.gather(c, ri)
            r_cols.append((_exchange(vals, back, rcv_back, g), c[1]))
        if outer and not b_cols:
            _, is_null = k.gather(b_keys[0], ri, want_valid=True)
            r_null = _exchange(is_null, back, rcv_back, g)
        if uniqueness == UNIQUE:
            # ---- owner, UNIQUE keys: a lhs row has at most one match, so the returned rows are put
            # back at their lhs positions (scatter) and, for INNER, the matched rows are compacted in
            # lhs order by the fused Filter kernel; no sort is needed
            n_l = lhs_keys[0][0].numel()
            flag = k.zeros(n_l, 6)
            k.scatter((k.zeros(origin.numel(), 6) + 1, 6), origin, flag)
            dense = []
            for t, dt in r_cols:
                d = k.zeros(n_l, dt)
                k.scatter((t, dt), origin, d)
                dense.append((d, dt))
            if outer:
                miss = k.zeros(n_l, 6)
                if r_null is not None:
                    k.scatter((r_null, 6), origin, miss)
                return k.iota(n_l), list(lhs_cols), dense, miss
            cols = k.compact(flag, [(k.iota(n_l), I64)] + list(lhs_cols) + dense)
            return cols[0][0], cols[1:1 + len(lhs_cols)], cols[1 + len(lhs_cols):], None
        # ---- owner: stable order by origin row
        order = k.order_by(origin)
        lhs_rows = k.gather((origin, I64), order)
        out_l = [(k.gather(c, lhs_rows), c[1]) for c in lhs_cols]
        out_r = [(k.gather(c, order), c[1]) for c in r_cols]
        if r_null is not None:
            r_null = k.gather((r_null, 6), order)
        return lhs_rows, out_l, out_r, r_null


# ---------------------------------------------------------------------------------------------
# Sort over row-range shards (sample sort)
# ---------------------------------------------------------------------------------------------
class ShardedSort(object):
    """Sort(order, project) over a table sharded by row range: every rank sorts its shard
    (ssb_sort_permutation, stable), splitters are taken from an all-gathered sample of the most
    significant key, rows are exchanged by key range with one all-to-all per column (NCCL), and
    every rank sorts what it received. After the call rank r holds the r-th range of the global
    order; rows with equal keys keep their global input order (all rows with one value of the most
    significant key meet on one rank, chunks arrive in rank order, and both sorts are stable), so
    the concatenation in rank order is the reference's output (cursor/core/sort.cc:150-322) whenever
    its order is total. Key and payload columns may carry is_null bytes (the same columns on every
    rank). Unsigned columns travel as their signed bit image."""

    SAMPLES_PER_RANK = 64

    def __init__(self, kernels, group=None):
        self.k, self.group = kernels, group

    def run(self, keys, descending, cols, key_nulls=None, col_nulls=None):
        """keys: [(tensor, SSB dtype)] most significant first; descending: [bool] per key;
        cols: payload columns; key_nulls / col_nulls: per column a uint8 tensor (1 = NULL) or None.
        Returns (sorted key columns, sorted payload columns) of this rank, followed by the two lists
        of is_null bytes when either was given. NULL keys sort first for ASCENDING and last for
        DESCENDING (sort.cc:174-238), so the rows with a NULL leading key all go to the first / last rank."""
        with_nulls = key_nulls is not None or col_nulls is not None
        key_nulls = list(key_nulls) if key_nulls is not None else [None] * len(keys)
        col_nulls = list(col_nulls) if col_nulls is not None else [None] * len(cols)
        with self.k.scope():
            out = self._run(keys, descending, cols, key_nulls, col_nulls)
        self.k.finish()
        return out if with_nulls else out[:2]

    def _run(self, keys, descending, cols, key_nulls, col_nulls):
        import torch
        import torch.distributed as dist
        k, g = self.k, self.group
        world = dist.get_world_size(g)
        device = keys[0][0].device
        # unsigned columns travel as their signed bit image: flipping the sign bit makes the image order-preserving
        flip = {3: -(1 << 63), 8: -(1 << 31)}.get(keys[0][1])   # UINT64, UINT32
        ordered = (lambda t: t) if flip is None else (lambda t: torch.bitwise_xor(t, torch.full((), flip, dtype=t.dtype, device=t.device)))
        BYTE = 6   # is_null bytes travel as BOOL columns

        def permute(perm, keys, cols, key_nulls, col_nulls):
            return ([(k.gather(c, perm), c[1]) for c in keys], [(k.gather(c, perm), c[1]) for c in cols],
                    [None if f is None else k.gather((f, BYTE), perm) for f in key_nulls],
                    [None if f is None else k.gather((f, BYTE), perm) for f in col_nulls])

        # 1. local stable sort
        perm = k.sort_perm(keys, descending, key_nulls)
        keys, cols, key_nulls, col_nulls = permute(perm, keys, cols, key_nulls, col_nulls)
        n = keys[0][0].numel()
        n_null = int(key_nulls[0].sum().item()) if key_nulls[0] is not None and n else 0
        # the rows whose leading key is not NULL: behind the NULLs (ASC) or in front of them (DESC)
        lo, hi = (0, n - n_null) if descending[0] else (n_null, n)
        first = keys[0][0][lo:hi]
        nn = hi - lo
        # 2. splitters from an all-gathered regular sample of the most significant key
        s = self.SAMPLES_PER_RANK
        if nn:
            pos = lo + (torch.arange(s, device=device, dtype=torch.int64) * nn) // s
            sample = k.gather(keys[0], pos)
        else:
            sample = first[:0]
        gathered = torch.cat(allgather_ragged(ordered(sample), g))
        gathered, _ = torch.sort(gathered)
        m = gathered.numel()
        if m and world > 1:
            splitters = gathered[[min(m - 1, (m * (r + 1)) // world) for r in range(world - 1)]]
            # 3. rows per destination: keys below the first splitter stay on rank 0, keys equal to a splitter go right of it
            asc = ordered(first if not descending[0] else first.flip(0))
            bounds = torch.searchsorted(asc.contiguous(), splitters, right=False).tolist()
        else:
            bounds = [0] * (world - 1)     # no row anywhere has a leading key that is not NULL
        edges = [0] + [int(b) for b in bounds] + [nn]
        counts = [edges[i + 1] - edges[i] for i in range(world)]
        if descending[0]:
            counts = counts[::-1]           # the local data is laid out largest first = destination 0 first
            counts[-1] += n_null            # ... and ends with the NULLs, which the last rank collects
        else:
            counts[0] += n_null             # NULLs come first and stay on the first rank
        recv = _exchange_counts(counts, device, g)
        # 4. exchange, 5. sort what arrived (chunks are sorted already; the stable sort merges them)
        ex = lambda t: _exchange(t, counts, recv, g)   # noqa: E731
        keys = [(ex(c[0]), c[1]) for c in keys]
        cols = [(ex(c[0]), c[1]) for c in cols]
        key_nulls = [None if f is None else ex(f) for f in key_nulls]
        col_nulls = [None if f is None else ex(f) for f in col_nulls]
        perm = k.sort_perm(keys, descending, key_nulls)
        return permute(perm, keys, cols, key_nulls, col_nulls)
