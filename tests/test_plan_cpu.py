"""ssb_program_plan: the expression compiler of libssb200.so without a device -- type checks, NULL propagation
(cross-checked against the result schemas the oracle binds for the same expressions), the shared-memory plan and
the error codes. Runs on the CPU."""
import numpy as np
import pytest

from supersonic_b200 import capi
from supersonic_b200 import ssplan as sp

I64, I32, F64, B = capi.INT64, capi.INT32, capi.DOUBLE, capi.BOOL
n = capi.node
INVALID_TYPE, INVALID_VALUE, NOT_IMPLEMENTED = 405, 407, 103


def _inputs(types):
    return [n(capi.OP_INPUT, t, [i]) for i, t in enumerate(types)]


def test_c2_plan_bytes_stages_and_budget(built):
    """BASELINE config 2: Filter(d < K, project e, Compute(e := a*b+c)): 32 B in, 8 B out per row."""
    nodes = _inputs([I64] * 4) + [n(capi.OP_MUL, I64, [0, 1]), n(capi.OP_ADD, I64, [4, 2]),
                                  n(capi.OP_CONST, I64, [], i64=1 << 19), n(capi.OP_LT, B, [3, 6])]
    for tile in (384, 768, 1024):
        for budget in (56 * 1024, 100 * 1024, 227 * 1024):
            p = capi.plan(nodes, [I64] * 4, [0] * 4, [5], predicate=7, tile=tile, smem_budget=budget)
            assert (p.bytes_per_input_row, p.bytes_per_output_row) == (32, 8)
            assert p.tile == tile and p.n_outputs == 1 and p.out_types[0] == I64 and p.out_nullable[0] == 0
            assert p.has_signaling == 0
            assert p.stages >= 2
            assert p.smem_bytes <= budget or p.stages == 2      # never below two stages: the budget then gives way
            assert p.smem_bytes >= p.stages * tile * 32          # every stage holds one tile of every input column
    small = capi.plan(nodes, [I64] * 4, [0] * 4, [5], predicate=7, tile=768, smem_budget=56 * 1024)
    large = capi.plan(nodes, [I64] * 4, [0] * 4, [5], predicate=7, tile=768, smem_budget=227 * 1024)
    assert large.stages >= small.stages and large.n_insn == small.n_insn


# (expression over columns a, b (NOT NULL) and na, nb (nullable) INT64, p / np BOOL; nodes; expected by the oracle)
def _nullability_cases():
    types = [I64, I64, I64, I64, B, B]           # a b na nb p np
    nullable = [0, 0, 1, 1, 0, 1]
    base = _inputs(types)
    c = []
    add = lambda text, extra, out_type: c.append((text, types, nullable, base + extra, out_type))   # noqa: E731
    add("(plus (col a) (col b))", [n(capi.OP_ADD, I64, [0, 1])], I64)
    add("(plus (col a) (col nb))", [n(capi.OP_ADD, I64, [0, 3])], I64)
    add("(less (col na) (col b))", [n(capi.OP_LT, B, [2, 1])], B)
    add("(is_null (col na))", [n(capi.OP_IS_NULL, B, [2])], B)
    add("(if_null (col na) (col b))", [n(capi.OP_IF_NULL, I64, [2, 1])], I64)
    add("(if_null (col na) (col nb))", [n(capi.OP_IF_NULL, I64, [2, 3])], I64)
    add("(divide_nulling (col a) (col b))", [n(capi.OP_CAST, F64, [0]), n(capi.OP_CAST, F64, [1]),
                                              n(capi.OP_DIV, F64, [6, 7], flags=capi.NODE_ZERO_NULLS)], F64)
    add("(divide_signaling (col a) (col b))", [n(capi.OP_CAST, F64, [0]), n(capi.OP_CAST, F64, [1]),
                                                n(capi.OP_DIV, F64, [6, 7], flags=capi.NODE_ZERO_FAILS)], F64)
    add("(modulus_nulling (col a) (col b))", [n(capi.OP_MOD, I64, [0, 1], flags=capi.NODE_ZERO_NULLS)], I64)
    add("(and (col p) (col np))", [n(capi.OP_AND, B, [4, 5])], B)
    add("(not (col np))", [n(capi.OP_NOT, B, [5])], B)
    add("(if (col p) (col a) (col nb))", [n(capi.OP_IF, I64, [4, 0, 3])], I64)
    add("(if (col np) (col a) (col b))", [n(capi.OP_IF, I64, [5, 0, 1])], I64)
    add("(nulling_if (col np) (col a) (col b))", [n(capi.OP_NULLING_IF, I64, [5, 0, 1])], I64)
    add("(nulling_if (col p) (col a) (col b))", [n(capi.OP_NULLING_IF, I64, [4, 0, 1])], I64)
    add("(negate (col na))", [n(capi.OP_NEGATE, I64, [2])], I64)
    return c


NULLABILITY = _nullability_cases()


@pytest.mark.parametrize("case", NULLABILITY, ids=[c[0] for c in NULLABILITY])
def test_null_propagation_matches_the_oracle_schema(ref, built, case):
    text, types, nullable, nodes, out_type = case
    info = capi.plan(nodes, types, nullable, [len(nodes) - 1])
    z = np.zeros(0, dtype=np.int64)
    zb = np.zeros(0, dtype=np.uint8)
    table = [sp.Column("a", sp.INT64, z), sp.Column("b", sp.INT64, z), sp.Column("na", sp.INT64, z, z.astype(bool)),
             sp.Column("nb", sp.INT64, z, z.astype(bool)), sp.Column("p", sp.BOOL, zb), sp.Column("np", sp.BOOL, zb, zb.astype(bool))]
    want = ref.run("(compute (as e %s) (scan 0))" % text, [table])
    assert want.code == 0, want.error
    assert info.n_outputs == 1 and info.out_types[0] == out_type == want.dtypes[0]
    assert bool(info.out_nullable[0]) == bool(want.nullable[0]), text
    assert bool(info.has_signaling) == ("signaling" in text)


def test_errors_carry_the_reference_codes(built):
    a = _inputs([I64, F64])
    with pytest.raises(capi.PlanError) as e:      # predicate must be BOOL
        capi.plan(a, [I64, F64], [0, 0], [0], predicate=0)
    assert e.value.code == INVALID_TYPE and "BOOL" in e.value.message
    with pytest.raises(capi.PlanError) as e:      # no implicit promotion at this level: the caller inserts CASTs
        capi.plan(a + [n(capi.OP_ADD, F64, [0, 1])], [I64, F64], [0, 0], [2])
    assert e.value.code == INVALID_TYPE
    with pytest.raises(capi.PlanError) as e:      # a node may only use earlier nodes
        capi.plan(a + [n(capi.OP_ADD, I64, [0, 3]), n(capi.OP_CONST, I64, [], i64=1)], [I64, F64], [0, 0], [2])
    assert e.value.code == INVALID_VALUE
    with pytest.raises(capi.PlanError) as e:      # input index out of range
        capi.plan([n(capi.OP_INPUT, I64, [5])], [I64], [0], [0])
    assert e.value.code == INVALID_VALUE
    with pytest.raises(capi.PlanError) as e:      # neither outputs nor predicate
        capi.plan(a, [I64, F64], [0, 0], [])
    assert e.value.code == INVALID_VALUE
    with pytest.raises(capi.PlanError) as e:      # unknown operator
        capi.plan(a + [n(999, I64, [0, 0])], [I64, F64], [0, 0], [2])
    assert e.value.code == NOT_IMPLEMENTED
    with pytest.raises(capi.PlanError) as e:      # more input columns than one kernel stages
        capi.plan(_inputs([I64] * 13), [I64] * 13, [0] * 13, [0])
    assert e.value.code == NOT_IMPLEMENTED
    with pytest.raises(capi.PlanError) as e:
        capi.plan(a, [I64, F64], [0, 0], [0], tile=100)
    assert e.value.code == INVALID_VALUE


def test_wide_plan_degrades_stage_by_stage(built):
    """Twelve 8-byte inputs and twelve outputs: a 768-row tile fits one SM with a single stage only (no overlap of
    copy and compute: ssb_program_create does not accept that and searches smaller tiles / column groups), 384 rows
    give two stages over the budget of four resident CTAs, 128 rows stay inside it; 1024 rows do not fit at all."""
    types = [I64] * 12
    nodes = _inputs(types) + [n(capi.OP_ADD, I64, [i, (i + 1) % 12]) for i in range(12)]
    outs = list(range(12, 24))
    budget = 56 * 1024
    p = capi.plan(nodes, types, [0] * 12, outs, tile=768, smem_budget=budget)
    assert p.stages == 1 and budget < p.smem_bytes <= 227 * 1024
    assert (p.bytes_per_input_row, p.bytes_per_output_row) == (96, 96)
    q = capi.plan(nodes, types, [0] * 12, outs, tile=384, smem_budget=budget)
    assert q.stages == 2 and budget < q.smem_bytes <= 227 * 1024
    r = capi.plan(nodes, types, [0] * 12, outs, tile=128, smem_budget=budget)
    assert r.stages >= 2 and r.smem_bytes <= budget
    with pytest.raises(capi.PlanError) as e:
        capi.plan(nodes, types, [0] * 12, outs, tile=1024, smem_budget=budget)
    assert e.value.code == NOT_IMPLEMENTED and "shared memory" in e.value.message
    # more budget buys stages, never a different program
    big = capi.plan(nodes, types, [0] * 12, outs, tile=384, smem_budget=227 * 1024)
    assert big.stages > q.stages and big.n_insn == q.n_insn


def test_group_aggregate_bind_time_memory_budget(ref, b200):
    """aggregate_groups.cc:465-469: the aggregator's first block (estimated_result_row_count rows) comes out of the
    operation's allocator at CreateCursor; an allocator that cannot hold it fails the bind with ERROR_MEMORY_EXCEEDED
    (aggregate_groups_test.cc:538-549). Checked without a GPU: SSPLAN_BIND_ONLY stops after CreateCursor."""
    table = [[sp.Column("col0", sp.INT32, [1, 3, 1, 3]), sp.Column("col1", sp.INT32, [3, -3, 4, -5])]]
    for plan, want in [("(group_opts none none 0 0 (named) (aggs (SUM col0 sum)) (scan 0))", 102),
                       ("(group_opts none none 0 0 (named col0) (aggs (SUM col1 sum)) (scan 0))", 102),
                       ("(group_opts none none 0 16 (named col0) (aggs (SUM col1 sum)) (scan 0))", 102),
                       ("(group_opts none none 0 4096 (named col0) (aggs (SUM col1 sum)) (scan 0))", 0),
                       ("(group_opts 1 1 0 none (named col0) (aggs (SUM col1 sum)) (scan 0))", 0),
                       ("(group_opts 1 1 1 none (named col0) (aggs (SUM col1 sum)) (scan 0))", 0)]:
        a = ref.run(plan, table, flags=sp.SSPLAN_BIND_ONLY)
        b = b200.run(plan, table, flags=sp.SSPLAN_BIND_ONLY)
        assert a.code == b.code == want, (plan, a.code, b.code)
