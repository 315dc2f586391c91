"""ctypes binding of the C ABI in include/supersonic_b200.h (libssb200.so).

Plumbing only: the hot path is the CUDA library. load() fails loudly when the library is
missing or does not export every entry point the header declares.
"""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "supersonic_b200", "lib", "libssb200.so")
HEADER = os.path.join(ROOT, "include", "supersonic_b200.h")

# SSB_* dtype numbers
INT32, INT64, UINT64, DATETIME, DOUBLE, BOOL, UINT32, FLOAT, DATE, ENUM = 1, 2, 3, 4, 5, 6, 8, 9, 10, 13
# expression ops (subset used from Python)
OP_INPUT, OP_CONST, OP_CAST = 1, 2, 3
OP_ADD, OP_SUB, OP_MUL, OP_DIV = 11, 12, 13, 14
OP_EQ, OP_NE, OP_LT, OP_LE, OP_GT, OP_GE = 20, 21, 22, 23, 24, 25
OP_AND, OP_OR, OP_NOT = 30, 31, 34
OP_NEGATE, OP_MOD, OP_IS_NULL, OP_IF_NULL, OP_IF, OP_NULLING_IF = 10, 15, 50, 51, 52, 53
NODE_NULL, NODE_ZERO_NULLS, NODE_ZERO_FAILS = 1, 2, 4
AGG_SUM, AGG_MIN, AGG_MAX, AGG_COUNT = 0, 1, 2, 3


class Column(C.Structure):
    _fields_ = [("data", C.c_void_p), ("nulls", C.c_void_p), ("dtype", C.c_int32), ("reserved", C.c_int32)]


class _Imm(C.Union):
    _fields_ = [("i64", C.c_int64), ("u64", C.c_uint64), ("f64", C.c_double), ("f32", C.c_float),
                ("i32", C.c_int32), ("u32", C.c_uint32), ("b", C.c_uint8)]


class ExprNode(C.Structure):
    _fields_ = [("op", C.c_int32), ("out_type", C.c_int32), ("arg", C.c_int32 * 3), ("flags", C.c_int32),
                ("imm", _Imm)]


class AggSpec(C.Structure):
    _fields_ = [("fn", C.c_int32), ("input", C.c_int32), ("in_type", C.c_int32), ("out_type", C.c_int32),
                ("in_nullable", C.c_int32), ("reserved", C.c_int32)]


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load():
    """Loads libssb200.so and verifies the exported entry points against the header."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError("libssb200.so is missing (%s): run __graft_entry__.build()" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    if missing:
        raise OSError("libssb200.so does not export: %s" % ", ".join(missing))
    P, I32, I64, U64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    sig = {
        "ssb_ctx_create": (C.c_int, [C.c_int, C.POINTER(P)]),
        "ssb_ctx_destroy": (None, [P]),
        "ssb_last_error": (C.c_char_p, [P]),
        "ssb_ctx_stream": (P, [P]),
        "ssb_ctx_sync": (C.c_int, [P]),
        "ssb_abi_version": (C.c_int, []),
        "ssb_ctx_launch_count": (I64, [P]),
        "ssb_ctx_last_kernel_ms": (C.c_int, [P, C.POINTER(C.c_float)]),
        "ssb_ctx_enable_timing": (C.c_int, [P, C.c_int]),
        "ssb_ctx_timer_start": (C.c_int, [P]),
        "ssb_ctx_timer_stop": (C.c_int, [P, C.POINTER(C.c_float)]),
        "ssb_malloc": (C.c_int, [P, C.c_size_t, C.POINTER(P)]),
        "ssb_free": (C.c_int, [P, P]),
        "ssb_malloc_host": (C.c_int, [P, C.c_size_t, C.POINTER(P)]),
        "ssb_free_host": (C.c_int, [P, P]),
        "ssb_memcpy_h2d": (C.c_int, [P, P, P, C.c_size_t]),
        "ssb_memcpy_d2h": (C.c_int, [P, P, P, C.c_size_t]),
        "ssb_transfer_bytes": (None, [C.POINTER(U64), C.POINTER(U64)]),
        "ssb_memcpy_d2d": (C.c_int, [P, P, P, C.c_size_t]),
        "ssb_memset": (C.c_int, [P, P, C.c_int, C.c_size_t]),
        "ssb_pointer_is_device": (C.c_int, [P]),
        "ssb_nulls_pack": (C.c_int, [P, P, I64, P]),
        "ssb_nulls_unpack": (C.c_int, [P, P, I64, P]),
        "ssb_generate": (C.c_int, [P, P, I64, I64, U64, U64, C.c_int, I64, U64]),
        "ssb_generate_host": (None, [P, I64, I64, U64, U64, C.c_int, I64, U64]),
        "ssb_program_create": (C.c_int, [P, C.POINTER(ExprNode), I32, I32, C.POINTER(I32), C.POINTER(I32),
                                         C.POINTER(I32), I32, I32, C.POINTER(P)]),
        "ssb_program_destroy": (None, [P]),
        "ssb_program_output_type": (I32, [P, I32]),
        "ssb_program_output_nullable": (I32, [P, I32]),
        "ssb_program_bytes_per_input_row": (I32, [P]),
        "ssb_program_bytes_per_output_row": (I32, [P]),
        "ssb_program_run": (C.c_int, [P, C.POINTER(Column), I64, C.POINTER(Column), P]),
        "ssb_program_run_sync": (C.c_int, [P, C.POINTER(Column), I64, C.POINTER(Column), C.POINTER(I64)]),
        "ssb_program_check_failure": (C.c_int, [P]),
        "ssb_group_create": (C.c_int, [P, I32, C.POINTER(I32), C.POINTER(I32), I32, C.POINTER(AggSpec), I64,
                                       C.POINTER(P)]),
        "ssb_group_destroy": (None, [P]),
        "ssb_group_update": (C.c_int, [P, C.POINTER(Column), C.POINTER(Column), I64]),
        "ssb_group_update_program": (C.c_int, [P, P, C.POINTER(Column), I64]),
        "ssb_jit_stats": (None, [C.POINTER(I64), C.POINTER(C.c_double), C.POINTER(I64)]),
        "ssb_jit_rows_compile": (C.c_int, [C.POINTER(ExprNode), C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                           C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(AggSpec),
                                           C.c_int32, C.c_int32, C.c_int32, C.c_char_p, I64, C.POINTER(I64)]),
        "ssb_group_finalize": (C.c_int, [P, C.POINTER(I64), C.POINTER(Column), C.POINTER(Column)]),
        "ssb_group_merge": (C.c_int, [P, I64, C.POINTER(Column), C.POINTER(Column)]),
        "ssb_join_build": (C.c_int, [P, I32, C.POINTER(Column), I64, I32, C.POINTER(P)]),
        "ssb_join_destroy": (None, [P]),
        "ssb_join_probe": (C.c_int, [P, C.POINTER(Column), I64, I32, C.POINTER(I64), C.POINTER(P), C.POINTER(P)]),
        "ssb_partition_rows": (C.c_int, [P, I32, C.POINTER(Column), I64, I32, I32, P, C.POINTER(I64)]),
        "ssb_gather": (C.c_int, [P, C.POINTER(Column), P, I64, C.POINTER(Column)]),
        "ssb_scatter": (C.c_int, [P, C.POINTER(Column), P, I64, C.POINTER(Column)]),
        "ssb_sort_permutation": (C.c_int, [P, I32, C.POINTER(Column), C.POINTER(I32), I64, P]),
        "ssb_program_plan": (C.c_int, [P, I32, I32, C.POINTER(I32), C.POINTER(I32), C.POINTER(I32), I32, I32, I32, C.c_uint32,
                                       P, C.c_char_p, I32]),
        "ssb_cluster_ids": (C.c_int, [P, I32, C.POINTER(Column), I64, P, P, C.POINTER(I64)]),
        "ssb_string_rank": (C.c_int, [P, P, P, I64, I64, P, P, C.POINTER(I64)]),
        "ssb_string_gather_offsets": (C.c_int, [P, P, P, I64, P, C.POINTER(I64)]),
        "ssb_string_gather_bytes": (C.c_int, [P, P, P, P, I64, P, P]),
        "ssb_string_shift_offsets": (C.c_int, [P, P, I64, I64, P]),
        "ssb_shard_join_build": (C.c_int, [P, C.POINTER(Column), I32, C.POINTER(Column), I64, C.POINTER(P)]),
        "ssb_shard_join_probe": (C.c_int, [P, C.POINTER(Column), I64, I32, C.POINTER(I64), C.POINTER(P), C.POINTER(P)]),
        "ssb_shard_join_payload": (C.c_int, [P, I32, C.POINTER(Column), C.POINTER(I64)]),
        "ssb_shard_join_form": (C.c_int, [P]),
        "ssb_shard_join_destroy": (None, [P]),
        "ssb_shard_join_probe_materialize": (C.c_int, [P, C.POINTER(Column), I64, I32, I32, C.POINTER(Column), I32, C.POINTER(I32),
                                                       C.POINTER(Column), P, C.POINTER(I64)]),
        "ssb_join_probe_materialize": (C.c_int, [P, C.POINTER(Column), I64, I32, I32, C.POINTER(Column), I32, C.POINTER(Column),
                                                 C.POINTER(Column), P, C.POINTER(I64)]),
        "ssb_join_table": (C.c_int, [P, C.POINTER(P), C.POINTER(I64)]),
        "ssb_join_attach_parts": (C.c_int, [P, I32, I32, C.POINTER(P), C.POINTER(I64), C.POINTER(I64), C.POINTER(P)]),
        "ssb_comm_unique_id": (C.c_int, [P]),
        "ssb_comm_create": (C.c_int, [P, P, I32, I32, C.POINTER(P)]),
        "ssb_comm_create_file": (C.c_int, [P, C.c_char_p, I32, I32, I32, C.POINTER(P)]),
        "ssb_comm_destroy": (None, [P]),
        "ssb_comm_rank": (I32, [P]),
        "ssb_comm_size": (I32, [P]),
        "ssb_comm_exchange_counts": (C.c_int, [P, C.POINTER(I64), C.POINTER(I64)]),
        "ssb_comm_all_to_all": (C.c_int, [P, I32, C.POINTER(P), C.POINTER(P), C.POINTER(I32), C.POINTER(I64), C.POINTER(I64)]),
        "ssb_shard_group_merge": (C.c_int, [P, P, C.POINTER(I64), C.POINTER(Column), C.POINTER(Column)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    _lib = lib
    return lib


class SsbError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "ssb error %d: %s" % (code, message))
        self.code = code


class Context(object):
    """One device context. Raises (never falls back) when no sm_100a device is present."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.ssb_ctx_create(device, C.byref(h))
        if rc != 0 or not h:
            raise SsbError(rc, "ssb_ctx_create failed: no usable B200 device %d (no CPU fallback exists)" % device)
        self.h = h

    def close(self):
        if self.h:
            self.lib.ssb_ctx_destroy(self.h)
            self.h = None

    def check(self, rc):
        if rc != 0:
            raise SsbError(rc, self.lib.ssb_last_error(self.h).decode(errors="replace"))

    def malloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.lib.ssb_malloc(self.h, nbytes, C.byref(p)))
        return p.value

    def free(self, ptr):
        self.check(self.lib.ssb_free(self.h, ptr))

    def malloc_host(self, nbytes):
        p = C.c_void_p()
        self.check(self.lib.ssb_malloc_host(self.h, nbytes, C.byref(p)))
        return p.value

    def free_host(self, ptr):
        self.check(self.lib.ssb_free_host(self.h, ptr))

    def sync(self):
        self.check(self.lib.ssb_ctx_sync(self.h))

    def h2d(self, dst, array):
        self.check(self.lib.ssb_memcpy_h2d(self.h, dst, array.ctypes.data, array.nbytes))
        self.sync()

    def d2h(self, array, src):
        self.check(self.lib.ssb_memcpy_d2h(self.h, array.ctypes.data, src, array.nbytes))
        self.sync()

    def generate(self, dst, rows, first_row, seed, stream, kind, lo, span):
        self.check(self.lib.ssb_generate(self.h, dst, rows, first_row, seed, stream, kind, lo, span))

    def launches(self):
        return self.lib.ssb_ctx_launch_count(self.h)

    def enable_timing(self, on=True):
        self.check(self.lib.ssb_ctx_enable_timing(self.h, 1 if on else 0))

    def timer_start(self):
        self.check(self.lib.ssb_ctx_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self.check(self.lib.ssb_ctx_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def last_kernel_ms(self):
        ms = C.c_float()
        self.check(self.lib.ssb_ctx_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value


def node(op, out_type, args=(), flags=0, i64=None, f64=None, b=None):
    n = ExprNode()
    n.op, n.out_type, n.flags = op, out_type, flags
    a = list(args) + [-1] * (3 - len(args))
    n.arg[0], n.arg[1], n.arg[2] = a
    if i64 is not None:
        n.imm.i64 = i64
    if f64 is not None:
        n.imm.f64 = f64
    if b is not None:
        n.imm.b = 1 if b else 0
    return n


class PlanInfo(C.Structure):
    _fields_ = [("tile", C.c_int32), ("stages", C.c_int32), ("smem_bytes", C.c_int32), ("n_insn", C.c_int32),
                ("n_tmp", C.c_int32), ("bytes_per_input_row", C.c_int32), ("bytes_per_output_row", C.c_int32),
                ("has_signaling", C.c_int32), ("n_outputs", C.c_int32), ("out_types", C.c_int32 * 16),
                ("out_nullable", C.c_int32 * 16)]


class PlanError(Exception):
    def __init__(self, code, message):
        Exception.__init__(self, "%d: %s" % (code, message))
        self.code, self.message = code, message


def plan(nodes, input_types, input_nullable, outputs, predicate=-1, tile=768, smem_budget=100 * 1024):
    """ssb_program_plan: validation and kernel planning of an expression DAG; needs no device."""
    lib = load()
    arr = (ExprNode * max(1, len(nodes)))(*nodes)
    it = (C.c_int32 * max(1, len(input_types)))(*input_types)
    inn = (C.c_int32 * max(1, len(input_types)))(*input_nullable)
    outs = (C.c_int32 * max(1, len(outputs)))(*outputs)
    info = PlanInfo()
    err = C.create_string_buffer(256)
    rc = lib.ssb_program_plan(arr, len(nodes), len(input_types), it, inn, outs, len(outputs), predicate, tile,
                              smem_budget, C.byref(info), err, 256)
    if rc != 0:
        raise PlanError(rc, err.value.decode())
    return info


class Program(object):
    def __init__(self, ctx, nodes, input_types, input_nullable, outputs, predicate=-1):
        self.ctx = ctx
        arr = (ExprNode * len(nodes))(*nodes)
        it = (C.c_int32 * max(1, len(input_types)))(*input_types)
        inn = (C.c_int32 * max(1, len(input_types)))(*input_nullable)
        outs = (C.c_int32 * max(1, len(outputs)))(*outputs)
        h = C.c_void_p()
        ctx.check(ctx.lib.ssb_program_create(ctx.h, arr, len(nodes), len(input_types), it, inn, outs,
                                             len(outputs), predicate, C.byref(h)))
        self.h = h
        self.n_in, self.n_out = len(input_types), len(outputs)

    def bytes_in(self):
        return self.ctx.lib.ssb_program_bytes_per_input_row(self.h)

    def bytes_out(self):
        return self.ctx.lib.ssb_program_bytes_per_output_row(self.h)

    def _cols(self, cols):
        arr = (Column * max(1, len(cols)))()
        for i, (data, nulls, dtype) in enumerate(cols):
            arr[i].data, arr[i].nulls, arr[i].dtype = data, nulls, dtype
        return arr

    def run(self, inputs, rows, outputs, d_out_rows=None):
        """Asynchronous launch. inputs/outputs: lists of (data_ptr, nulls_ptr_or_None, dtype)."""
        self.ctx.check(self.ctx.lib.ssb_program_run(self.h, self._cols(inputs), rows, self._cols(outputs), d_out_rows))

    def run_sync(self, inputs, rows, outputs):
        n = C.c_int64()
        self.ctx.check(self.ctx.lib.ssb_program_run_sync(self.h, self._cols(inputs), rows, self._cols(outputs),
                                                         C.byref(n)))
        return n.value

    def close(self):
        if self.h:
            self.ctx.lib.ssb_program_destroy(self.h)
            self.h = None


COMM_ID_BYTES = 128


class Comm(object):
    """A communicator of libssb200.so (NCCL bound at run time) over the context's stream: one per rank.
    `unique_id()` on rank 0, hand the bytes to the other ranks by any channel, then Comm(ctx, id, world, rank)."""

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        rc = load().ssb_comm_unique_id(buf)
        if rc != 0:
            raise SsbError(rc, "ssb_comm_unique_id failed (is NCCL loadable?)")
        return bytes(buf)

    def __init__(self, ctx, unique_id, world, rank):
        self.ctx, self.world, self.rank = ctx, world, rank
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        h = C.c_void_p()
        ctx.check(ctx.lib.ssb_comm_create(ctx.h, buf, world, rank, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.ctx.lib.ssb_comm_destroy(self.h)
            self.h = None
