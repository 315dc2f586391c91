"""The run-time compiled aggregation kernel (csrc/jit.cu, csrc/jit_rows.h) without a GPU: NVRTC turns the
plan-specialised source into an sm_100a cubin here (nvcc's runtime sibling needs no device), so a plan whose
generated code does not compile fails on the CPU suite, not on the GPU box. Values are checked on the GPU by
tests/test_device_gpu.py::test_fused_* in their `jit` mode."""
import ctypes as C
import os

import pytest

from supersonic_b200 import capi


def _compile(nodes, types, nullable, outputs, predicate, n_keys, aggs, groups, threads=0, rows_per_thread=0):
    lib = capi.load()
    arr = (capi.ExprNode * len(nodes))(*nodes)
    specs = (capi.AggSpec * len(aggs))()
    for i, (fn, inp, it, ot, inn) in enumerate(aggs):
        specs[i].fn, specs[i].input, specs[i].in_type, specs[i].out_type, specs[i].in_nullable = fn, inp, it, ot, inn
    text = C.create_string_buffer(1 << 20)
    size = C.c_int64()
    rc = lib.ssb_jit_rows_compile(arr, len(nodes), len(types), (C.c_int32 * len(types))(*types), (C.c_int32 * len(types))(*nullable),
                                  (C.c_int32 * len(outputs))(*outputs), len(outputs), predicate, n_keys, len(aggs), specs,
                                  groups, threads, rows_per_thread, text, len(text), C.byref(size))
    return rc, size.value, text.value.decode()


def _q1():
    n = capi.node
    F64, I64, B = capi.DOUBLE, capi.INT64, capi.BOOL
    types = [F64, F64, F64, F64, I64, I64, I64]
    nodes = [n(capi.OP_INPUT, t, [j]) for j, t in enumerate(types)]
    nodes += [n(capi.OP_CONST, F64, [], f64=1.0), n(capi.OP_SUB, F64, [7, 2]), n(capi.OP_MUL, F64, [1, 8]),
              n(capi.OP_ADD, F64, [7, 3]), n(capi.OP_MUL, F64, [9, 10]), n(capi.OP_CONST, I64, [], i64=2450),
              n(capi.OP_LE, B, [6, 12])]
    aggs = [(capi.AGG_SUM, i, F64, F64, 0) for i in range(5)] + [(capi.AGG_COUNT, -1, I64, capi.UINT64, 0)]
    return nodes, types, [0] * 7, [4, 5, 0, 1, 9, 11, 2], 13, 2, aggs


def test_q1_plan_compiles_to_an_sm100a_cubin(built):
    rc, size, text = _compile(*_q1(), groups=6)
    assert rc == 0, text[-4000:]
    assert size > 10000
    # the generated prelude: one X-macro line per input column, instruction and aggregate
    assert "enum { T = 192, R = 1, G = 6" in text and "N_IN = 7" in text and "NK = 2, A = 6" in text
    assert text.count("\n  X(") == 7 + 6 + text.split("#define SSB_JIT_PROGRAM(X)")[1].split("#define")[0].count("\n  X(")
    assert '#include "jit_rows.h"' in text


@pytest.mark.parametrize("threads,rows_per_thread", [(256, 2), (128, 4), (64, 1)])
def test_launch_shapes_compile(built, threads, rows_per_thread):
    rc, size, text = _compile(*_q1(), groups=8, threads=threads, rows_per_thread=rows_per_thread)
    assert rc == 0 and size > 0, text[-4000:]
    assert "T = %d, R = %d, G = 8" % (threads, rows_per_thread) in text


def test_nullable_mixed_type_plan_compiles(built):
    """NULL-carrying inputs, a cast, MIN / MAX / COUNT(column) / SUM(INT32): every accumulate form of jit_rows.h."""
    n = capi.node
    F64, I64, I32, B = capi.DOUBLE, capi.INT64, capi.INT32, capi.BOOL
    types = [I64, I64, F64, I32]
    nodes = [n(capi.OP_INPUT, t, [j]) for j, t in enumerate(types)]
    nodes += [n(capi.OP_CONST, I64, [], i64=3), n(capi.OP_MUL, I64, [1, 4]), n(capi.OP_CONST, I64, [], i64=-2500),
              n(capi.OP_GT, B, [5, 6]), n(capi.OP_CAST, I64, [3]), n(capi.OP_ADD, I64, [5, 8])]
    aggs = [(capi.AGG_SUM, 0, I64, I64, 1), (capi.AGG_MIN, 1, F64, F64, 1), (capi.AGG_MAX, 1, F64, F64, 1),
            (capi.AGG_COUNT, 1, F64, capi.UINT64, 1), (capi.AGG_COUNT, -1, I64, capi.UINT64, 0), (capi.AGG_SUM, 2, I32, I32, 0)]
    rc, size, text = _compile(nodes, types, [1, 1, 1, 0], [0, 9, 2, 3], 7, 1, aggs, groups=8)
    assert rc == 0 and size > 0, text[-4000:]


def test_scalar_aggregate_and_signaling_division_compile(built):
    n = capi.node
    I64, B = capi.INT64, capi.BOOL
    nodes = [n(capi.OP_INPUT, I64, [0]), n(capi.OP_INPUT, I64, [1]), n(capi.OP_CONST, I64, [], i64=10), n(capi.OP_LT, B, [1, 2]),
             n(capi.OP_DIV, I64, [0, 1], flags=capi.NODE_ZERO_FAILS)]
    aggs = [(capi.AGG_SUM, 0, I64, I64, 0), (capi.AGG_MAX, 0, I64, I64, 0), (capi.AGG_COUNT, -1, I64, capi.UINT64, 0)]
    rc, size, text = _compile(nodes, [I64, I64], [0, 0], [4], 3, 0, aggs, groups=1)
    assert rc == 0 and size > 0, text[-4000:]


def test_plans_outside_the_kernel_limits_are_refused(built):
    nodes, types, nullable, outputs, predicate, n_keys, aggs = _q1()
    rc, size, text = _compile(nodes, types, nullable, outputs, predicate, n_keys, aggs, groups=9)
    assert rc == 103 and size == 0 and "limits" in text


def test_many_groups_form_compiles(built):
    """groups = 0: no CTA-local entries, every row goes to the global table (csrc/jit_rows.h, Spec::G == 0)."""
    rc, size, text = _compile(*_q1(), groups=0)
    assert rc == 0 and size > 0, text[-4000:]
    assert "G = 0" in text


_T = {"i32": capi.INT32, "i64": capi.INT64, "u32": capi.UINT32, "u64": capi.UINT64, "f32": capi.FLOAT, "f64": capi.DOUBLE}


@pytest.mark.parametrize("tname", sorted(_T))
def test_every_operator_family_compiles_for_a_type(built, tname):
    """One plan per operand type that strings together every operator family alu() folds (arithmetic incl. signaling /
    nulling division and modulus, comparisons, three-valued logic, IF / NULLING_IF / IFNULL / IS_NULL, casts, and for the
    integer types the bitwise operators and shifts), over nullable inputs, aggregated with SUM / MIN / MAX / COUNT: NVRTC
    must accept the specialised source for every one of them (a plan that failed to compile would silently keep the
    interpreting kernels on the GPU)."""
    n = capi.node
    t, B, I64, F64 = _T[tname], capi.BOOL, capi.INT64, capi.DOUBLE
    integer = tname[0] in "iu"
    nodes = [n(capi.OP_INPUT, t, [0]), n(capi.OP_INPUT, t, [1]), n(capi.OP_INPUT, I64, [2])]                 # 0 a, 1 b, 2 key
    nodes += [n(capi.OP_ADD, t, [0, 1]), n(capi.OP_SUB, t, [3, 1]), n(capi.OP_MUL, t, [4, 0]),                # 3 4 5
              n(capi.OP_DIV, t, [5, 1], flags=capi.NODE_ZERO_NULLS), n(capi.OP_MOD, I64 if not integer else t, [5, 1], flags=capi.NODE_ZERO_NULLS),   # 6 7
              n(capi.OP_LT, B, [0, 1]), n(capi.OP_EQ, B, [0, 1]), n(capi.OP_GE, B, [3, 1]),                   # 8 9 10
              n(capi.OP_AND, B, [8, 9]), n(capi.OP_OR, B, [11, 10]), n(capi.OP_NOT, B, [12]),                  # 11 12 13
              n(capi.OP_IF, t, [13, 0, 1]), n(capi.OP_NULLING_IF, t, [8, 14, 6]), n(capi.OP_IF_NULL, t, [15, 0]),   # 14 15 16
              n(capi.OP_IS_NULL, B, [6]), n(capi.OP_CAST, F64, [16]), n(capi.OP_CAST, I64, [18]),               # 17 18 19
              n(capi.OP_NEGATE, I64 if tname in ("u32", "u64") else t, [0])]                                    # 20
    outs = [2, 16, 18, 19, 20, 7]
    if integer:
        nodes += [n(40, t, [0, 1]), n(41, t, [21, 1]), n(42, t, [22, 0]), n(44, t, [23]), n(45, t, [24, 1]), n(46, t, [25, 1])]   # 21..26 bitwise, shifts
        outs.append(26)
    aggs = [(capi.AGG_SUM, 0, t, t, 1), (capi.AGG_MIN, 1, F64, F64, 1), (capi.AGG_MAX, 2, I64, I64, 1),
            (capi.AGG_COUNT, 3, nodes[20].out_type, capi.UINT64, 1), (capi.AGG_SUM, 4, nodes[7].out_type, nodes[7].out_type, 1),
            (capi.AGG_COUNT, -1, I64, capi.UINT64, 0)]
    if integer:
        aggs.append((capi.AGG_MAX, 5, t, t, 1))
    for groups in (0, 3):
        rc, size, text = _compile(nodes, [t, t, I64], [1, 1, 0], outs, 17, 1, aggs, groups=groups)
        assert rc == 0 and size > 0, text[-6000:]


def test_q1_kernel_resources_at_the_default_launch_shape(built, tmp_path, monkeypatch):
    """The shape bench.py's Q1 leg runs (192 threads x 1 row, prefetch 1, 3 CTAs per SM): at most 84 registers per thread
    (three CTAs of 192 threads must fit the register file) and no spills beyond the out-of-line cold paths' frames. Two rows
    per thread with prefetch spilled and lost a factor of two (profiles/r2h_q1_jit.txt): this pins the compiled shape."""
    import re
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump on this box")
    path = str(tmp_path / "q1.cubin")
    monkeypatch.setenv("SSB200_JIT_DUMP", path)
    for name in ("SSB200_JIT_THREADS", "SSB200_JIT_ROWS", "SSB200_JIT_PREFETCH", "SSB200_JIT_MIN_CTAS"):
        monkeypatch.delenv(name, raising=False)
    rc, size, text = _compile(*_q1(), groups=6)
    assert rc == 0 and os.path.getsize(path) == size, text[-2000:]
    assert "enum { T = 192, R = 1, G = 6, MIN_CTAS = 3, PREFETCH = 1" in text
    usage = subprocess.run([cuobjdump, "-res-usage", path], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    m = re.search(r"Function ssb_jit_rows:\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", usage)
    assert m, usage[-2000:]
    regs, stack, shared, local = map(int, m.groups())
    assert regs <= 84 and stack <= 64 and local == 0, usage[-1000:]
