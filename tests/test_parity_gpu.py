"""Parity of the CUDA path with the oracle (the unmodified reference), through the same
plan text compiled against both supersonic.h implementations. Integer, bool, index and row
order results must be bit-exact; DOUBLE sums are bit-exact on exactly-summable payloads and
within 1 ulp per addend otherwise (tolerance stated at the test)."""
import itertools

import numpy as np
import pytest

from cases import GOLDEN, check_result, same_results
from supersonic_b200 import ssplan as sp

pytestmark = pytest.mark.gpu

NP = {sp.INT32: np.int32, sp.INT64: np.int64, sp.UINT32: np.uint32, sp.UINT64: np.uint64,
      sp.FLOAT: np.float32, sp.DOUBLE: np.float64, sp.BOOL: np.bool_, sp.DATE: np.int32,
      sp.DATETIME: np.int64}
TNAME = {sp.INT32: "i32", sp.INT64: "i64", sp.UINT32: "u32", sp.UINT64: "u64", sp.FLOAT: "f32",
         sp.DOUBLE: "f64", sp.BOOL: "b", sp.DATE: "d", sp.DATETIME: "dt"}


def random_column(rng, name, dtype, n, nullable, small=False):
    if dtype == sp.BOOL:
        data = rng.integers(0, 2, n).astype(np.bool_)
    elif dtype in (sp.FLOAT, sp.DOUBLE):
        data = (rng.integers(-1000, 1000, n) / 8.0).astype(NP[dtype])
    else:
        info = np.iinfo(NP[dtype])
        lo, hi = (max(info.min, -50), min(info.max, 50)) if small else (max(info.min, -2**31), min(info.max, 2**31 - 1))
        data = rng.integers(lo, hi, n, dtype=np.int64).astype(NP[dtype])
        # sprinkle zeros and extremes
        data[rng.integers(0, n, max(1, n // 50))] = 0
    nulls = (rng.random(n) < 0.1) if nullable else None
    return sp.Column(name, dtype, data, is_null=nulls)


def table(rng, n, small=False):
    cols = []
    for dt, nm in TNAME.items():
        cols.append(random_column(rng, nm, dt, n, False, small))
        cols.append(random_column(rng, "n" + nm, dt, n, True, small))
    return cols


@pytest.mark.parametrize("case", GOLDEN, ids=[c[0] for c in GOLDEN])
@pytest.mark.parametrize("next_rows", [0, 1, 3])
def test_reference_vectors(b200, case, next_rows):
    _, plan, tables, expected, ordered = case
    check_result(b200.run(plan, tables, next_max_rows=next_rows), expected, ordered)


ARITH = ["plus", "minus", "multiply", "divide_quiet", "divide_nulling", "cpp_divide_nulling", "modulus_nulling"]
CMP = ["equal", "not_equal", "less", "less_or_equal", "greater", "greater_or_equal"]
NUMS = ["i32", "i64", "u32", "u64", "f32", "f64"]
INTS = ["i32", "i64", "u32", "u64"]


@pytest.mark.parametrize("n", [1, 31, 1024, 5000, 100003])
def test_expression_matrix(ref, b200, n):
    rng = np.random.default_rng(n)
    cols = table(rng, n, small=True)
    exprs = []
    for op in ARITH:
        for x, y in itertools.product(NUMS, NUMS):
            if op.startswith("modulus") and ("f" in x or "f" in y):
                continue
            exprs.append("(%s (col %s) (col n%s))" % (op, x, y))
    for op in CMP:
        for x, y in itertools.product(NUMS + ["b", "d", "dt"], NUMS + ["b", "d", "dt"]):
            exprs.append("(%s (col n%s) (col %s))" % (op, x, y))
    for op in ["and", "or", "and_not", "xor"]:
        exprs.append("(%s (col nb) (less (col ni32) (col i64)))" % op)
        exprs.append("(%s (col b) (col nb))" % op)
    for op in ["bitwise_and", "bitwise_or", "bitwise_xor", "bitwise_and_not"]:
        for x, y in itertools.product(INTS, INTS):
            exprs.append("(%s (col %s) (col n%s))" % (op, x, y))
    for x in NUMS:
        exprs.append("(negate (col n%s))" % x)
        exprs.append("(is_null (col n%s))" % x)
        exprs.append("(if_null (col n%s) (col %s))" % (x, x))
        exprs.append("(if (col nb) (col %s) (col n%s))" % (x, x))
        exprs.append("(nulling_if (col nb) (col %s) (col n%s))" % (x, x))
        for t in ["INT64", "DOUBLE", "FLOAT", "UINT64", "INT32", "UINT32"]:
            exprs.append("(cast %s (col n%s))" % (t, x))
    for x in INTS:
        exprs.append("(is_odd (col n%s))" % x)
        exprs.append("(is_even (col %s))" % x)
        exprs.append("(bitwise_not (col n%s))" % x)
        exprs.append("(shift_left (col %s) (i32 3))" % x)
        exprs.append("(shift_right (col n%s) (i32 2))" % x)
    exprs += ["(case (col ni32) (col f64) (i32 1) (col ni64) (i32 3) (col i64))",
              "(case (col i64) (null INT64) (col ni32) (col i64) (col u32) (col ni64))",
              "(in (col ni32) (col i32) (i32 2) (col ni64))", "(in (col f64) (col ni32) (f64 1.5))",
              "(in (col i32) (i32 1) (null INT32))"]
    # multiply-add superinstruction: two roundings for DOUBLE (no FMA contraction), wrap for ints
    exprs += ["(plus (multiply (col f64) (col f64)) (col f64))", "(plus (multiply (col i64) (col i64)) (col i64))",
              "(plus (multiply (col i32) (col i32)) (col i32))", "(plus (col f64) (multiply (col f64) (f64 1.1)))",
              "(plus (multiply (plus (col f64) (f64 0.3)) (col f64)) (col f64))",
              "(minus (i64 5) (col i64))", "(greater (col i64) (i64 3))", "(less_or_equal (i32 7) (col i32))"]
    exprs += ["(not (col nb))", "(cast DATETIME (col nd))",
              "(plus (multiply (col i64) (col i32)) (minus (col f64) (col nu32)))",
              "(if (less (col i32) (i32 0)) (negate (col i32)) (col i32))",
              "(and (less (col i32) (col i64)) (or (is_null (col nf64)) (greater (col nf64) (f64 1.5))))"]
    # evaluate in groups through one Compute each (also exercises multi-output programs)
    bad = []
    for i in range(0, len(exprs), 6):
        chunk = exprs[i:i + 6]
        plan = "(compute (compound %s) (scan 0))" % " ".join("(as c%d %s)" % (k, e) for k, e in enumerate(chunk))
        a = ref.run(plan, [cols])
        b = b200.run(plan, [cols])
        try:
            same_results(a, b)
        except AssertionError as err:
            # narrow down to the single expression
            for e in chunk:
                p1 = "(compute (as c %s) (scan 0))" % e
                try:
                    same_results(ref.run(p1, [cols]), b200.run(p1, [cols]))
                except AssertionError as err1:
                    bad.append((e, str(err1)[:200]))
            if not bad:
                bad.append((chunk, str(err)[:200]))
    assert not bad, bad[:10]


def test_signaling_division_fails(ref, b200):
    cols = [sp.Column("a", sp.INT32, [1, 2, 3]), sp.Column("b", sp.INT32, [1, 0, 3])]
    for op in ["divide_signaling", "cpp_divide_signaling", "modulus_signaling"]:
        plan = "(compute (%s (col a) (col b)) (scan 0))" % op
        a, b = ref.run(plan, [cols]), b200.run(plan, [cols])
        assert a.code == 104 and b.code == 104, (op, a.code, b.code, b.error)
    ok = [sp.Column("a", sp.INT32, [1, 2, 3]), sp.Column("b", sp.INT32, [1, 2, 3])]
    same_results(ref.run("(compute (cpp_divide_signaling (col a) (col b)) (scan 0))", [ok]),
                 b200.run("(compute (cpp_divide_signaling (col a) (col b)) (scan 0))", [ok]))


from cases import EVALUATE_CASES, SIGNALING_CASES, signaling_tables  # noqa: E402


@pytest.mark.parametrize("case", EVALUATE_CASES, ids=[c[0] for c in EVALUATE_CASES])
def test_bound_expression_tree_evaluate(ref, b200, case):
    """Expression::Bind + BoundExpressionTree::Evaluate of the mirror (test/guide/primer.cc's
    entry point) against the reference's golden values and against the oracle."""
    _, plan, tables, expected, code = case
    want, got = ref.run(plan, tables), b200.run(plan, tables)
    assert want.code == code and got.code == code, (want.code, got.code, got.error)
    if code == 0:
        check_result(got, expected, True)
        same_results(want, got)


def test_evaluate_random_expressions(ref, b200):
    rng = np.random.default_rng(12)
    cols = table(rng, 5000, small=True)
    for e in ["(plus (multiply (col i64) (col i32)) (col ni64))", "(if (col nb) (col f64) (negate (col nf64)))",
              "(compound (as x (less (col ni32) (col i64))) (as y (cast DOUBLE (col u32))) (col nu64))"]:
        for cap in (0, 1024, 4999):
            plan = "(evaluate %s 0 %d)" % (e, cap)
            same_results(ref.run(plan, [cols]), b200.run(plan, [cols]))


@pytest.mark.parametrize("case", SIGNALING_CASES, ids=[c[0] for c in SIGNALING_CASES])
def test_signaling_ops_follow_skip_vectors(ref, b200, case):
    """SURVEY 8a9 / ADVICE r1: a signaling division under IF / CASE / AND / OR / IFNULL, next to a NULL
    operand, or above a Filter fails exactly when the reference's skip-vector evaluation fails, and gives
    the reference's values otherwise (the kernel evaluates every row; the rows that count travel as a guard)."""
    _, plan, code = case
    tables = signaling_tables()
    want, got = ref.run(plan, tables), b200.run(plan, tables)
    assert want.code == code and got.code == code, (want.code, got.code, got.error)
    if code == 0:
        same_results(want, got, ordered="group" not in plan, sort_cols=[0] if "group" in plan else None)


@pytest.mark.parametrize("n,sel", [(10_000_000, 2**19), (1_000_003, 2**10), (2049, 2**20), (1024, 0)])
def test_filter_project_c1(ref, b200, n, sel):
    """BASELINE config 1: Compute(a*b+c) then Filter(d<K) over 4 x INT64, bit-exact, in order."""
    rng = np.random.default_rng(42)
    cols = [sp.Column("a", sp.INT64, rng.integers(-2**31, 2**31, n)),
            sp.Column("b", sp.INT64, rng.integers(-2**31, 2**31, n)),
            sp.Column("c", sp.INT64, rng.integers(-2**62, 2**62, n)),
            sp.Column("d", sp.INT64, rng.integers(0, 2**20, n))]
    plan = ("(filter (less (col d) (i64 %d)) (named e) (compute (compound (as e (plus (multiply "
            "(col a) (col b)) (col c))) (col d)) (scan 0)))" % sel)
    same_results(ref.run(plan, [cols], next_max_rows=16384), b200.run(plan, [cols], next_max_rows=16384))
    plan_b = ("(filter (less (col d) (i64 %d)) (all) (compute (compound (as e (plus (multiply "
              "(col a) (col b)) (col c))) (col a) (col b) (col c) (col d)) (scan 0)))" % sel)
    same_results(ref.run(plan_b, [cols], next_max_rows=16384), b200.run(plan_b, [cols], next_max_rows=16384))


@pytest.mark.parametrize("n", [777, 100_000])
def test_filter_comparison_predicates(ref, b200, n):
    """Every fast comparison form feeding the compaction directly (compare fused with the
    predicate step), next to predicates that go through a materialised BOOL."""
    rng = np.random.default_rng(n)
    cols = table(rng, n, small=True)
    preds = ["(less (col i64) (i64 2))", "(greater_or_equal (col i64) (col i64))", "(greater (col i32) (i32 0))",
             "(less_or_equal (col f64) (f64 0.5))", "(equal (col i32) (i32 1))", "(not_equal (col i64) (i64 0))",
             "(less (plus (col i64) (i64 1)) (col i64))", "(greater (multiply (col f64) (f64 2.0)) (col f64))",
             "(less (col ni64) (i64 2))", "(and (less (col i32) (i32 3)) (greater (col f64) (f64 -1.0)))",
             "(less (i64 1) (col i64))", "(col b)", "(less (col i32) (col i64))"]
    for pr in preds:
        plan = ("(project (named e i32 nf64) (filter %s (all) (compute (compound (col i64) (col i32) (col f64) "
                "(col ni64) (col nf64) (col b) (as e (plus (multiply (col i64) (col i64)) (col i64)))) (scan 0))))" % pr)
        a = ref.run(plan, [cols])
        b = b200.run(plan, [cols])
        assert a.code == 0, (pr, a.error)
        try:
            same_results(a, b)
        except AssertionError as e:
            raise AssertionError("%s: %s" % (pr, e))


def test_filter_with_nulls_and_stacked_operators(ref, b200):
    rng = np.random.default_rng(3)
    cols = table(rng, 70001)
    plans = [
        "(filter (less (col ni32) (col i64)) (named ni64 f64 nb) (scan 0))",
        "(filter (col nb) (all) (scan 0))",
        "(project (named x) (filter (greater (col x) (i64 0)) (all) (compute (compound (as x (plus (col i64) (col ni32))) (col b)) (scan 0))))",
        "(filter (col b) (named y) (compute (as y (multiply (col x) (col x))) (filter (less (col x) (f64 100)) (named x b) (compute (compound (as x (col f64)) (col b)) (scan 0)))))",
        "(compute (plus (col s) (i64 1)) (group (named b) (aggs (SUM i64 s)) (scan 0)))",
    ]
    for plan in plans:
        same_results(ref.run(plan, [cols]), b200.run(plan, [cols]), ordered="group" not in plan)


@pytest.mark.parametrize("n,groups", [(1_000_000, 1000), (2_000_000, 300_000), (100_000, 3), (5000, 5000)])
def test_group_aggregate(ref, b200, n, groups):
    """C3 shape: SUM(DOUBLE) + COUNT per INT64 key. The payload k * 2^-10 (k < 2^20) makes every
    partial sum exactly representable, so the result is bit-exact in any order."""
    rng = np.random.default_rng(n)
    cols = [sp.Column("k", sp.INT64, rng.integers(0, groups, n)),
            sp.Column("v", sp.DOUBLE, rng.integers(0, 2**20, n) / 1024.0),
            sp.Column("w", sp.INT32, rng.integers(-1000, 1000, n), is_null=rng.random(n) < 0.2),
            sp.Column("u", sp.UINT64, rng.integers(0, 2**40, n))]
    plan = ("(group (named k) (aggs (SUM v sum_v) (COUNT \"\" cnt) (MIN w mn) (MAX w mx) (SUM w sw INT64) "
            "(COUNT w cw) (MAX u mu) (MIN v mv)) (scan 0))")
    same_results(ref.run(plan, [cols]), b200.run(plan, [cols]), ordered=False, sort_cols=[0])


def test_group_multi_key_and_null_keys(ref, b200):
    rng = np.random.default_rng(11)
    n = 300_000
    cols = [sp.Column("a", sp.INT32, rng.integers(0, 50, n), is_null=rng.random(n) < 0.05),
            sp.Column("b", sp.INT64, rng.integers(-3, 3, n)),
            sp.Column("c", sp.BOOL, rng.integers(0, 2, n).astype(np.bool_), is_null=rng.random(n) < 0.3),
            sp.Column("v", sp.INT64, rng.integers(-10**6, 10**6, n))]
    plan = "(group (named a b c) (aggs (SUM v s) (COUNT \"\" n) (MIN v mn)) (scan 0))"
    same_results(ref.run(plan, [cols]), b200.run(plan, [cols]), ordered=False, sort_cols=[0, 1, 2])
    plan1 = "(group (named a) (aggs (SUM v s) (MAX b m)) (scan 0))"
    same_results(ref.run(plan1, [cols]), b200.run(plan1, [cols]), ordered=False, sort_cols=[0])


def test_group_double_sum_tolerance(ref, b200):
    """Arbitrary doubles: the GPU adds in a different order than the reference's sequential
    loop (column_aggregator.cc:112-123). Tolerance: |diff| <= n_group * ulp(max partial)."""
    rng = np.random.default_rng(5)
    n = 200_000
    cols = [sp.Column("k", sp.INT64, rng.integers(0, 100, n)), sp.Column("v", sp.DOUBLE, rng.random(n))]
    plan = "(group (named k) (aggs (SUM v s)) (scan 0))"
    a, b = ref.run(plan, [cols]), b200.run(plan, [cols])
    oa, ob = np.argsort(a.columns[0]), np.argsort(b.columns[0])
    assert np.array_equal(a.columns[0][oa], b.columns[0][ob])
    sa, sb = a.columns[1][oa], b.columns[1][ob]
    per_group = n / 100
    assert np.all(np.abs(sa - sb) <= per_group * np.spacing(np.maximum(sa, sb)))


@pytest.mark.parametrize("uniq", ["UNIQUE", "NOT_UNIQUE"])
@pytest.mark.parametrize("jt", ["INNER", "LEFT_OUTER"])
def test_hash_join(ref, b200, uniq, jt):
    """C4 shape at test size: probe fk uniform over build keys, order-preserving output."""
    rng = np.random.default_rng(17)
    nb, npr = 100_000, 700_001
    pk = rng.permutation(nb).astype(np.int64) * 3   # gaps: some probes miss
    if uniq == "NOT_UNIQUE":
        pk[rng.integers(0, nb, nb // 10)] = pk[rng.integers(0, nb, nb // 10)]
    build = [sp.Column("pk", sp.INT64, pk), sp.Column("payload", sp.INT64, rng.integers(0, 10**9, nb)),
             sp.Column("pn", sp.DOUBLE, rng.random(nb), is_null=rng.random(nb) < 0.1)]
    probe = [sp.Column("fk", sp.INT64, rng.integers(0, nb * 3, npr), is_null=rng.random(npr) < 0.02),
             sp.Column("lv", sp.INT64, rng.integers(0, 10**9, npr))]
    plan = ("(hash_join %s (named fk) (named pk) (multi (0 (all)) (1 (named payload pn))) %s (scan 0) (scan 1))"
            % (jt, uniq))
    same_results(ref.run(plan, [probe, build], next_max_rows=8192), b200.run(plan, [probe, build], next_max_rows=8192))


def test_hash_join_big_fanout(ref, b200):
    """hash_join_test.cc:321-353: 5 keys x 1100 duplicates, more result rows per key than a block."""
    lhs = [sp.Column("k", sp.INT32, np.arange(5))]
    rhs = [sp.Column("k2", sp.INT32, np.tile(np.arange(5), 1100)), sp.Column("w", sp.INT64, np.arange(5500))]
    plan = "(hash_join INNER (named k) (named k2) (multi (0 (all)) (1 (named w))) NOT_UNIQUE (scan 0) (scan 1))"
    same_results(ref.run(plan, [lhs, rhs]), b200.run(plan, [lhs, rhs]))


@pytest.mark.parametrize("n", [1, 2, 1000, 250_000])
def test_sort(ref, b200, n):
    rng = np.random.default_rng(n)
    cols = [sp.Column("a", sp.INT64, rng.integers(-50, 50, n), is_null=rng.random(n) < 0.1),
            sp.Column("b", sp.DOUBLE, rng.integers(-1000, 1000, n) / 4.0),
            sp.Column("c", sp.INT32, rng.integers(-2**31, 2**31 - 1, n)),
            sp.Column("id", sp.INT64, np.arange(n))]
    # a full key (ending in the unique id) makes the order total, so the unstable reference sort
    # and the stable GPU sort must agree exactly
    for order in ["(a ASC) (b DESC) (id ASC)", "(c DESC) (id ASC)", "(a DESC) (c ASC) (id DESC)", "(b ASC) (id ASC)"]:
        plan = "(sort (order %s) (all) (scan 0))" % order
        same_results(ref.run(plan, [cols]), b200.run(plan, [cols]))


def test_q1_shape(ref, b200):
    """C5 shape: Filter(ship <= D) -> Compute(disc_price, charge) -> GroupAggregate({rf, ls})."""
    rng = np.random.default_rng(23)
    n = 600_000
    cols = [sp.Column("qty", sp.DOUBLE, rng.integers(1, 51, n).astype(np.float64)),
            sp.Column("price", sp.DOUBLE, rng.integers(100, 10000, n) / 4.0),
            sp.Column("disc", sp.DOUBLE, rng.integers(0, 5, n) / 16.0),
            sp.Column("tax", sp.DOUBLE, rng.integers(0, 5, n) / 16.0),
            sp.Column("rf", sp.INT64, rng.integers(0, 3, n)), sp.Column("ls", sp.INT64, rng.integers(0, 2, n)),
            sp.Column("ship", sp.INT64, rng.integers(0, 2500, n))]
    plan = ("(group (named rf ls) (aggs (SUM qty sum_qty) (SUM price sum_price) (SUM disc_price sum_disc_price) "
            "(SUM charge sum_charge) (SUM disc sum_disc) (COUNT \"\" cnt)) "
            "(compute (compound (col rf) (col ls) (col qty) (col price) (col disc) "
            "(as disc_price (multiply (col price) (minus (f64 1) (col disc)))) "
            "(as charge (multiply (multiply (col price) (minus (f64 1) (col disc))) (plus (f64 1) (col tax))))) "
            "(filter (less_or_equal (col ship) (i64 2450)) (all) (scan 0))))")
    # dyadic payloads: products and sums stay exactly representable -> bit-exact in any order
    same_results(ref.run(plan, [cols]), b200.run(plan, [cols]), ordered=False, sort_cols=[0, 1])


from cases import BOUND_PAIRS, bound_tables  # noqa: E402


@pytest.mark.parametrize("pair", BOUND_PAIRS, ids=[p[0] for p in BOUND_PAIRS])
def test_bound_factories(ref, b200, pair):
    """BoundCompute / BoundFilter / BoundProject / BoundScanView / BoundGroupAggregate /
    BoundScalarAggregate / BoundSort of the mirror against the reference's."""
    name, unbound, bound, ordered = pair
    tables = bound_tables()
    want, got = ref.run(bound, tables), b200.run(bound, tables)
    assert want.code == 0 and got.code == 0, (want.error, got.error)
    assert want.names == got.names and want.dtypes == got.dtypes
    same_results(want, got, ordered=ordered, sort_cols=None if ordered else list(range(len(want.columns))))


def test_bound_filter_rejects_non_bool_predicate(ref, b200):
    tables = bound_tables()
    plan = "(bound_filter (plus (col a) (col b)) (all) (bound_scan 0))"
    want, got = ref.run(plan, tables), b200.run(plan, tables)
    assert want.code != 0 and got.code == want.code


@pytest.mark.parametrize("n,groups", [(5000, 7), (300_000, 1000), (300_000, 5)])
def test_group_first_last(ref, b200, n, groups):
    """FIRST / LAST = the first / last non-NULL input of the group in input order
    (column_aggregator.cc:108-166); 300k rows cross the slices the GPU path feeds separately."""
    rng = np.random.default_rng(n + groups)
    cols = [sp.Column("k", sp.INT64, rng.integers(0, groups, n)),
            sp.Column("a", sp.INT64, rng.integers(-10**9, 10**9, n), is_null=rng.random(n) < 0.3),
            sp.Column("x", sp.DOUBLE, rng.random(n)),
            sp.Column("c", sp.INT32, rng.integers(-1000, 1000, n).astype(np.int32), is_null=rng.random(n) < 0.9)]
    plan = ("(group (named k) (aggs (FIRST a fa) (LAST a la) (FIRST x fx) (LAST x lx) (FIRST c fc) (LAST c lc) "
            "(SUM x sx) (COUNT a ca)) (scan 0))")
    want, got = ref.run(plan, [cols]), b200.run(plan, [cols])
    assert want.code == 0 and got.code == 0, (want.error, got.error)
    # SUM(x) of uniform doubles reorders on the GPU: compare it separately within a tolerance
    sx = want.names.index("sx")
    ow, og = np.argsort(want.columns[0]), np.argsort(got.columns[0])
    assert np.allclose(want.columns[sx][ow], got.columns[sx][og], rtol=1e-12)
    for r in (want, got):
        r.columns[sx] = np.zeros_like(r.columns[sx])
    same_results(want, got, ordered=False, sort_cols=[0])
    plan2 = "(scalar_agg (aggs (FIRST a fa) (LAST a la) (LAST c lc) (COUNT \"\" n)) (scan 0))"
    same_results(ref.run(plan2, [cols]), b200.run(plan2, [cols]))


@pytest.mark.parametrize("n", [1000, 200_003])
def test_wide_plans_run_as_column_groups(ref, b200, n):
    """BASELINE config 2 variant B (ProjectAllAttributes: the eight columns plus e = a*b+c under the
    filter) and other wide plans: the library evaluates them as column groups that share the
    predicate; rows, order, NULLs and values must equal the reference's single pass."""
    rng = np.random.default_rng(n)
    cols = [sp.Column("a", sp.INT64, rng.integers(-2**31, 2**31, n)), sp.Column("b", sp.INT64, rng.integers(-2**31, 2**31, n)),
            sp.Column("c", sp.INT64, rng.integers(-2**62, 2**62, n)), sp.Column("d", sp.INT64, rng.integers(0, 2**20, n)),
            sp.Column("e0", sp.INT64, rng.integers(0, 100, n), is_null=rng.random(n) < 0.2),
            sp.Column("f", sp.DOUBLE, rng.random(n)), sp.Column("g", sp.INT32, rng.integers(-5, 5, n).astype(np.int32)),
            sp.Column("h", sp.INT64, rng.integers(0, 10**12, n), is_null=rng.random(n) < 0.5)]
    variant_b = ("(filter (less (col d) (i64 524288)) (all) (compute (compound (as e (plus (multiply (col a) (col b)) (col c))) "
                 "(col a) (col b) (col c) (col d) (col e0) (col f) (col g) (col h)) (scan 0)))")
    same_results(ref.run(variant_b, [cols]), b200.run(variant_b, [cols]))
    wide_compute = ("(compute (compound (as o1 (plus (col a) (col b))) (as o2 (minus (col c) (col d))) (as o3 (multiply (col f) (f64 2))) "
                    "(as o4 (plus (col e0) (col h))) (as o5 (cast INT64 (col g))) (as o6 (is_null (col h))) (col a) (col b) (col c) (col d) "
                    "(as o7 (if_null (col e0) (i64 -1))) (as o8 (less (col f) (f64 0.5)))) (scan 0))")
    same_results(ref.run(wide_compute, [cols]), b200.run(wide_compute, [cols]))
    nullable_pred = ("(filter (greater (col h) (i64 500000000000)) (all) (compute (compound (col a) (col b) (col c) (col d) (col e0) (col f) "
                     "(col g) (col h) (as s (plus (col e0) (col h)))) (scan 0)))")
    same_results(ref.run(nullable_pred, [cols]), b200.run(nullable_pred, [cols]))


def test_group_over_filter_streams_host_chunks(ref, b200, monkeypatch):
    """GroupAggregate over Filter/Compute of a host table is fed to the GPU chunk by chunk (bounded
    device memory); with 4096-row chunks the 100k-row table takes 25 calls into one hash table,
    and FIRST / LAST must still see the rows in input order."""
    monkeypatch.setenv("SSB200_GROUP_CHUNK_ROWS", "4096")
    rng = np.random.default_rng(77)
    n = 100_003
    cols = [sp.Column("k", sp.INT64, rng.integers(0, 300, n), is_null=rng.random(n) < 0.01),
            sp.Column("a", sp.INT64, rng.integers(-10**6, 10**6, n), is_null=rng.random(n) < 0.2),
            sp.Column("b", sp.INT64, rng.integers(0, 100, n))]
    plan = ("(group (named k) (aggs (SUM s ss) (MIN a mn) (COUNT a ca) (COUNT \"\" n)) "
            "(filter (less (col b) (i64 70)) (all) (compute (compound (col k) (col a) (col b) (as s (plus (col a) (col b)))) (scan 0))))")
    same_results(ref.run(plan, [cols]), b200.run(plan, [cols]), ordered=False, sort_cols=[0])
    plan_fl = ("(group (named k) (aggs (FIRST s fs) (LAST s ls)) "
               "(compute (compound (col k) (as s (plus (col a) (col b)))) (scan 0)))")
    same_results(ref.run(plan_fl, [cols]), b200.run(plan_fl, [cols]), ordered=False, sort_cols=[0])
    plan_scalar = "(scalar_agg (aggs (SUM s ss) (COUNT \"\" n)) (compute (as s (multiply (col b) (i64 3))) (scan 0)))"
    same_results(ref.run(plan_scalar, [cols]), b200.run(plan_scalar, [cols]))


@pytest.mark.parametrize("narrow", ["1", "0"])
def test_streaming_with_transfer_narrowing(ref, b200, monkeypatch, narrow):
    """Host tables stream to the GPU in chunks; 64-bit integer columns whose chunk fits 32 bits
    travel narrowed and are widened by the kernel. Columns that fit in some chunks only (the
    program variant changes from chunk to chunk), never, always, NULL cells with wide garbage,
    DATETIME: the result must not depend on it."""
    monkeypatch.setenv("SSB200_CHUNK_ROWS", "8192")
    monkeypatch.setenv("SSB200_NARROW_TRANSFERS", narrow)
    rng = np.random.default_rng(5)
    n = 100_000
    sometimes = rng.integers(-1000, 1000, n)
    sometimes[30_000:50_000] = rng.integers(-2**62, 2**62, 20_000)       # these chunks do not fit
    sometimes[77_777] = 2**31                                             # one value just outside int32
    edge = rng.integers(-2**31, 2**31, n)
    edge[0], edge[1] = -2**31, 2**31 - 1                                  # the int32 limits fit
    garbage = rng.integers(0, 100, n)
    gnull = rng.random(n) < 0.3
    garbage[gnull] = 2**50                                                # wide garbage under NULL
    cols = [sp.Column("s", sp.INT64, sometimes), sp.Column("e", sp.INT64, edge),
            sp.Column("w", sp.INT64, rng.integers(-2**62, 2**62, n)), sp.Column("g", sp.INT64, garbage, is_null=gnull),
            sp.Column("t", sp.DATETIME, rng.integers(0, 10**9, n)), sp.Column("d", sp.INT64, rng.integers(0, 2**20, n))]
    plan = ("(filter (less (col d) (i64 600000)) (all) (compute (compound (as x (plus (multiply (col s) (col e)) (col w))) "
            "(col s) (col e) (col g) (col t) (as y (plus (col g) (col d)))) (scan 0)))")
    same_results(ref.run(plan, [cols], next_max_rows=1024), b200.run(plan, [cols], next_max_rows=5000))


# ------------------------------------------------------------------------------------------------
# Added after the last GPU run of round 1 (oracle-pinned on the CPU): kept at the end of the GPU suite so that
# everything measured before still runs first.
@pytest.mark.parametrize("jt", ["RIGHT_OUTER", "FULL_OUTER"])
def test_unsupported_join_types_fail_at_the_first_lookup(ref, b200, jt):
    """hash_join.cc:713-726: any join type binds; the other two are refused once the probe side has produced a row."""
    plan = "(hash_join %s (named k) (named k) (multi (0 (named v)) (1 (rename (v w)))) UNIQUE (scan 0) (scan 1))" % jt
    t = lambda n: [sp.Column("k", sp.INT64, np.arange(n)), sp.Column("v", sp.INT64, np.arange(n))]   # noqa: E731
    for nl, nr in ((3, 3), (3, 0)):
        a, b = ref.run(plan, [t(nl), t(nr)]), b200.run(plan, [t(nl), t(nr)])
        assert a.code == 103 and b.code == 103, (nl, nr, a.code, b.code, b.error)
        assert jt in b.error


from cases import GOLDEN_LATE  # noqa: E402

# least new machinery first (key images, then the sort cursor's row limit, then the selection cursor)
_LATE_ORDER = ["sort_signed_zero", "group_signed_zero_keys", "join_signed_zero_keys", "extended_sort"]
LATE = sorted(GOLDEN_LATE, key=lambda c: next((i for i, p in enumerate(_LATE_ORDER) if c[0].startswith(p)), len(_LATE_ORDER)))


@pytest.mark.parametrize("next_rows", [0, 1, 3])
@pytest.mark.parametrize("case", LATE, ids=[c[0] for c in LATE])
def test_reference_vectors_late(b200, case, next_rows):
    _, plan, tables, expected, ordered = case
    check_result(b200.run(plan, tables, next_max_rows=next_rows), expected, ordered)



def test_nan_join_keys_never_match(ref, b200):
    """row_hash_set.cc:487-498 confirms a hash hit with operator==, which no NaN satisfies: a NaN key finds nothing
    and is found by nothing. The kernels treat a NaN key like a NULL key."""
    nan = float("nan")
    build = [sp.Column("pk", sp.DOUBLE, [1.0, nan, 2.0, nan, 3.0, 2.0]), sp.Column("w", sp.INT64, [10, 20, 30, 40, 50, 60])]
    probe = [sp.Column("fk", sp.DOUBLE, [nan, 1.0, 3.0, nan, 7.0, 2.0]), sp.Column("lv", sp.INT64, [1, 2, 3, 4, 5, 6])]
    for jt in ("INNER", "LEFT_OUTER"):
        plan = "(hash_join %s (named fk) (named pk) (multi (0 (all)) (1 (named w))) NOT_UNIQUE (scan 0) (scan 1))" % jt
        same_results(ref.run(plan, [probe, build]), b200.run(plan, [probe, build]))
    rng = np.random.default_rng(8)
    n = 50_000
    pk = rng.integers(0, 3000, n).astype(np.float32)
    pk[rng.random(n) < 0.05] = np.nan
    fk = rng.integers(0, 3000, 3 * n).astype(np.float32)
    fk[rng.random(3 * n) < 0.05] = np.nan
    big_b = [sp.Column("pk", sp.FLOAT, pk), sp.Column("w", sp.INT64, np.arange(n))]
    big_p = [sp.Column("fk", sp.FLOAT, fk), sp.Column("lv", sp.INT64, np.arange(3 * n))]
    plan = "(hash_join LEFT_OUTER (named fk) (named pk) (multi (0 (all)) (1 (named w))) NOT_UNIQUE (scan 0) (scan 1))"
    same_results(ref.run(plan, [big_p, big_b], next_max_rows=8192), b200.run(plan, [big_p, big_b], next_max_rows=8192))


def test_nan_group_keys_are_refused(ref, b200):
    """The reference makes every row with a NaN key a group of its own (operator== after the hash); the GPU table
    compares bit images and would merge them, so the plan is refused with ERROR_NOT_IMPLEMENTED instead of being
    answered differently. Float keys without NaN (and NaN hidden under NULL) aggregate as before."""
    nan = float("nan")
    cols = [sp.Column("k", sp.DOUBLE, [1.0, nan, 1.0, nan, 2.0]), sp.Column("v", sp.INT64, [1, 2, 3, 4, 5])]
    plan = "(group (named k) (aggs (SUM v s) (COUNT \"\" n)) (scan 0))"
    want, got = ref.run(plan, [cols]), b200.run(plan, [cols])
    assert want.code == 0 and want.rows == 4          # two NaN rows, two groups
    assert got.code == 103 and "NaN" in got.error
    plan_f = "(group (named k) (aggs (SUM v s)) (compute (compound (col k) (col v)) (filter (greater (col v) (i64 0)) (all) (scan 0))))"
    assert b200.run(plan_f, [cols]).code == 103        # the fused child path checks as well
    ok = [sp.Column("k", sp.DOUBLE, [1.0, nan, 1.0, 0.5, 2.0], is_null=[False, True, False, False, False]),
          sp.Column("v", sp.INT64, [1, 2, 3, 4, 5])]
    same_results(ref.run(plan, [ok]), b200.run(plan, [ok]), ordered=False, sort_cols=[0])


def test_apply_to_children_with_a_pass_through_transformer(ref, b200):
    """Cursor::ApplyToChildren (cursor/base/cursor.h:210) the way the reference's tests use it with their spy cursors
    (e.g. aggregate_clusters_test.cc:84-103): every child of the root cursor is wrapped by a CursorTransformer, then
    the root; results must not change. A wrapped child is no longer a GPU cursor, so its rows reach the operator
    through Next() -- the seam OperationTest-style harnesses rely on. Row-wise chains over a scan are one fused cursor
    here and have no child cursors; the others must hand over every child."""
    rng = np.random.default_rng(21)
    n = 5000
    t = [sp.Column("k", sp.INT32, rng.integers(0, 50, n).astype(np.int32)), sp.Column("v", sp.INT64, rng.integers(-99, 99, n)),
         sp.Column("d", sp.DOUBLE, rng.integers(0, 64, n) / 8.0, is_null=rng.random(n) < 0.1), sp.Column("id", sp.INT64, np.arange(n))]
    u = [sp.Column("pk", sp.INT32, np.arange(50, dtype=np.int32)), sp.Column("w", sp.INT64, np.arange(50) * 7)]
    plans = [("(group (named k) (aggs (SUM v s) (MAX d m) (COUNT \"\" c)) (scan 0))", 1, False),
             ("(group (named k) (aggs (SUM v s)) (filter (greater (col v) (i64 0)) (all) (scan 0)))", 1, False),
             ("(sort (order (k DESC) (id ASC)) (all) (compute (compound (col k) (col id) (as e (plus (col v) (i64 1)))) (scan 0)))", 1, True),
             ("(hash_join LEFT_OUTER (named k) (named pk) (multi (0 (named id v)) (1 (named w))) UNIQUE (scan 0) (scan 1))", 2, True),
             ("(merge_union_all (order (id ASC)) (scan 0) (scan 0))", 1, True),   # here: a sort over ONE concatenating cursor
             ("(aggregate_clusters (named k) (aggs (SUM v s)) (sort (order (k ASC) (id ASC)) (all) (scan 0)))", 1, True),
             ("(compute (as e (multiply (col v) (col v))) (sort (order (id DESC)) (all) (scan 0)))", 1, True),
             ("(filter (less (col v) (i64 0)) (all) (scan 0))", 0, True)]
    for plan, children, ordered in plans:
        want = ref.run(plan, [t, u], flags=sp.SSPLAN_SPY)
        got = b200.run(plan, [t, u], flags=sp.SSPLAN_SPY)
        assert got.spied_children == children, (plan, got.spied_children)
        same_results(want, got, ordered=ordered, sort_cols=[0])
        same_results(b200.run(plan, [t, u]), got, ordered=ordered, sort_cols=[0])


def test_bound_expression_factories_and_do_evaluate(ref, b200):
    """Expressions assembled bottom-up from the BOUND factories (BoundNamedAttribute, BoundConst*, BoundPlus, BoundLess,
    BoundIf, BoundCastTo, BoundAlias, BoundCompoundExpression: expression/core/*_bound_expressions.h) and wrapped with
    CreateBoundExpressionTree give the same values as the reference's, through BoundCompute / BoundFilter cursors and
    through the virtual BoundExpression::DoEvaluate(view, skip vectors) (expression/base/expression.h:46-93)."""
    rng = np.random.default_rng(33)
    n = 1000
    t = [[sp.Column("a", sp.INT32, rng.integers(-50, 50, n).astype(np.int32)), sp.Column("b", sp.INT64, rng.integers(-10**6, 10**6, n), is_null=rng.random(n) < 0.2),
          sp.Column("x", sp.DOUBLE, rng.integers(-64, 64, n) / 8.0), sp.Column("u", sp.UINT32, rng.integers(0, 2**32, n).astype(np.uint32)),
          sp.Column("f", sp.BOOL, rng.integers(0, 2, n).astype(np.bool_), is_null=rng.random(n) < 0.1), sp.Column("s", sp.STRING, [b"ab", b"b", b""][0:1] * n)]]
    exprs = ["(compound (as e (plus (multiply (col a) (col b)) (i64 7))) (less (col a) (i32 2)) (if_null (col b) (i64 -1)))",
             "(compound (divide_nulling (col x) (cast DOUBLE (col a))) (cpp_divide_nulling (col u) (col a)) (modulus_nulling (col b) (i64 7)))",
             "(if (and (col f) (greater (col x) (f64 0))) (cast INT64 (col a)) (col b))",
             "(compound (nulling_if (col f) (col x) (negate (col x))) (is_null (col b)) (not (col f)) (xor (col f) (is_odd (col a))))",
             "(compound (bitwise_and (col u) (u32 255)) (shift_left (col a) (i32 3)) (equal (col u) (col a)) (greater_or_equal (col b) (col a)))",
             "(if (equal (col a) (i32 0)) (i32 0) (cpp_divide_signaling (i32 100) (col a)))"]
    for e in exprs:
        same_results(ref.run("(bound_bx_compute %s (bound_scan 0))" % e, t), b200.run("(bound_bx_compute %s (bound_scan 0))" % e, t))
        same_results(ref.run("(bx_evaluate %s 0)" % e, t), b200.run("(bx_evaluate %s 0)" % e, t))
    same_results(ref.run("(bound_bx_filter (or (less (col x) (f64 -2)) (is_null (col b))) (named a b x) (bound_scan 0))", t),
                 b200.run("(bound_bx_filter (or (less (col x) (f64 -2)) (is_null (col b))) (named a b x) (bound_scan 0))", t))
    same_results(ref.run('(bound_bx_filter (equal (col s) (str "ab")) (named s a) (bound_scan 0))', t),
                 b200.run('(bound_bx_filter (equal (col s) (str "ab")) (named s a) (bound_scan 0))', t))
    fails = "(bx_evaluate (cpp_divide_signaling (col a) (minus (col a) (col a))) 0)"
    assert ref.run(fails, t).code == b200.run(fails, t).code == sp.ERROR_EVALUATION_ERROR


# GroupAggregate's memory contract (aggregate_groups.cc:452-480; aggregate_groups_test.cc:538-573, 849-870): the result
# block starts with estimated_result_row_count rows under a soft quota and fails with ERROR_MEMORY_EXCEEDED when it cannot
# grow; BestEffortGroupAggregate never fails for lack of memory (it may emit a key in several rows).
BUDGET_TABLE = [[sp.Column("col0", sp.INT32, [1, 3, 1, 3]), sp.Column("col1", sp.INT32, [3, -3, 4, -5])]]
BUDGET_PLANS = [
    "(group_opts 1 1 0 none (named col0) (aggs (SUM col1 sum)) (scan 0))",       # :551-573 quota for one row, two groups
    "(group_opts none 2 0 none (named col0) (aggs (SUM col1 sum)) (scan 0))",    # :575-600 the block grows
    "(group_opts 64 1 0 none (named col0) (aggs (SUM col1 sum)) (scan 0))",      # quota for seven rows
    "(group_opts none none 0 0 (named) (aggs (SUM col0 sum)) (scan 0))",         # :538-549 MemoryLimit(0): bind fails
    "(group_opts none none 0 16 (named col0) (aggs (SUM col1 sum)) (scan 0))",   # allocator too small for the first block
    "(group_opts none none 0 4096 (named col0) (aggs (SUM col1 sum)) (scan 0))",
]


@pytest.mark.parametrize("plan", BUDGET_PLANS)
def test_group_aggregate_memory_budget_matches_the_reference(ref, b200, plan):
    a, b = ref.run(plan, BUDGET_TABLE), b200.run(plan, BUDGET_TABLE)
    assert a.code == b.code, (plan, a.code, b.code, b.error)
    if a.code == 0:
        same_results(a, b, ordered=False)
    else:
        assert a.code == 102


def test_best_effort_group_aggregate_never_runs_out_of_memory(ref, b200):
    """With room for one result row the reference emits partial results (every input row on its own here); the GPU
    cursor aggregates in HBM and emits every key once. Both are valid best-effort outputs: re-aggregated they agree."""
    plan = "(group_opts 1 1 1 none (named col0) (aggs (SUM col1 sum)) (scan 0))"
    a, b = ref.run(plan, BUDGET_TABLE), b200.run(plan, BUDGET_TABLE)
    assert a.code == 0 and b.code == 0, (a.code, b.code, b.error)

    def totals(r):
        out = {}
        for k, v in zip(r.columns[0], r.columns[1]):
            out[int(k)] = out.get(int(k), 0) + int(v)
        return out
    assert totals(a) == totals(b) == {1: 7, 3: -8}
    assert b.rows == 2


def test_cursor_trees_over_a_file_scan_match_the_reference(ref, b200, tmp_path):
    """rows arrive from outside the process in the reference's block format (file_io.cc:70-420). The file
    is written by the REFERENCE's FileOutput; FileInput (a CPU cursor of <= 8192-row chunks) then feeds the GPU cursors
    through Next(), and the results must equal the reference's over the same file."""
    rng = np.random.default_rng(7)
    rows = 30000
    words = ["ab", "", "Supersonic", "B200", "columnar"]
    table = [[sp.Column("k", sp.INT64, rng.integers(0, 50, rows)),
              sp.Column("a", sp.INT64, rng.integers(-1000, 1000, rows), is_null=rng.random(rows) < 0.1),
              sp.Column("x", sp.DOUBLE, rng.integers(0, 1 << 16, rows) / 4.0),
              sp.Column("s", sp.STRING, [words[i] for i in rng.integers(0, len(words), rows)])]]
    path = str(tmp_path / "rows.ssb")
    assert ref.run("(file_write %s (scan 0))" % path, table).code == 0
    src = "(bound_file_read %s 0)" % path
    plans = [("(bound_filter (less (col a) (i64 10)) (named k x s) %s)" % src, True),
             ("(bound_compute (compound (as e (plus (col a) (col k))) (col s)) %s)" % src, True),
             ("(bound_group (named k) (aggs (SUM x sx) (COUNT a ca) (MIN a mn) (COUNT \"\" n)) %s)" % src, False),
             ("(bound_group (named s) (aggs (SUM x sx) (MAX a mx)) %s)" % src, False),
             ("(bound_sort (order (s ASC) (k DESC) (x ASC) (a ASC)) (all) %s)" % src, True)]
    for plan, ordered in plans:
        same_results(ref.run(plan, table), b200.run(plan, table), ordered=ordered)
    # and back out: the GPU result written by the mirror's FileOutput holds the rows the reference writes for its own
    # result (the chunking follows the size of the views Next() returns, which is not part of the data: the reference
    # reads both files back)
    fa, fb = str(tmp_path / "a.ssb"), str(tmp_path / "b.ssb")
    plan = "(sort (order (k ASC) (x ASC) (a ASC) (s ASC)) (all) (filter (less (col a) (i64 10)) (all) (scan 0)))"
    ra, rb = ref.run("(file_write %s %s)" % (fa, plan), table), b200.run("(file_write %s %s)" % (fb, plan), table)
    assert ra.code == 0 and rb.code == 0, (ra.code, rb.code, rb.error)
    same_results(ra, rb)
    same_results(ref.run("(file_read %s 0)" % fa, table), ref.run("(file_read %s 0)" % fb, table))


@pytest.mark.parametrize("groups", [7, 3000])
def test_distinct_aggregates_match_the_reference(ref, b200, groups):
    """COUNT / SUM DISTINCT (column_aggregator.cc:333-433) beside plain aggregates, over nullable inputs, NULL keys, two
    key columns, DOUBLE and STRING inputs, under a Filter + Compute child, and as a ScalarAggregate: the passes of
    GroupCursor::RunDistinct against the reference's per-group hash sets."""
    rng = np.random.default_rng(groups)
    rows = 40000
    words = ["ab", "", "Supersonic", "B200", "columnar", "x"]
    t = [[sp.Column("k", sp.INT64, rng.integers(0, groups, rows), is_null=rng.random(rows) < 0.03),
          sp.Column("k2", sp.INT32, rng.integers(0, 3, rows)),
          sp.Column("v", sp.INT64, rng.integers(-20, 20, rows), is_null=rng.random(rows) < 0.1),
          sp.Column("w", sp.DOUBLE, rng.integers(0, 16, rows) / 4.0),
          sp.Column("s", sp.STRING, [words[i] for i in rng.integers(0, len(words), rows)], is_null=rng.random(rows) < 0.05),
          sp.Column("u", sp.INT32, rng.integers(0, 1000, rows))]]
    plans = [
        "(group (named k) (aggs (distinct COUNT v cv) (distinct SUM v sv) (SUM v s) (COUNT \"\" n) (MIN w mw)) (scan 0))",
        "(group (named k k2) (aggs (distinct SUM w sw) (distinct COUNT v cv) (distinct COUNT s cs) (MAX u mu)) (scan 0))",
        "(group (named k2) (aggs (distinct COUNT u cu) (distinct SUM u su)) (scan 0))",
        "(scalar_agg (aggs (distinct COUNT v cv) (distinct SUM w sw) (distinct COUNT s cs) (COUNT \"\" n)) (scan 0))",
        "(group (named k2) (aggs (distinct COUNT e ce) (SUM e se)) (compute (compound (col k2) (as e (plus (col v) (col u)))) "
        "(filter (less (col u) (i32 500)) (all) (scan 0))))",
        "(group (named k) (aggs (distinct COUNT v cv)) (scan 0))",
    ]
    for plan in plans:
        same_results(ref.run(plan, t), b200.run(plan, t), ordered=False)


@pytest.mark.parametrize("quota", [64, 1 << 30])
def test_hybrid_group_aggregate_matches_the_reference(ref, b200, quota):
    """aggregate.h:309-336 HybridGroupAggregate: exact aggregation, DISTINCT included, whatever the memory quota (the
    reference spills sorted runs to temporary files; here the groups live in HBM)."""
    rng = np.random.default_rng(quota % 1000)
    rows = 20000
    t = [[sp.Column("k", sp.INT64, rng.integers(0, 500, rows), is_null=rng.random(rows) < 0.03),
          sp.Column("v", sp.INT64, rng.integers(-20, 20, rows), is_null=rng.random(rows) < 0.1),
          sp.Column("w", sp.DOUBLE, rng.integers(0, 16, rows) / 4.0)]]
    for plan in ["(hybrid_group %d (named k) (aggs (distinct COUNT v c) (SUM v s) (distinct SUM w sw) (COUNT \"\" n)) (scan 0))" % quota,
                 "(hybrid_group %d (named k) (aggs (SUM v s) (MIN w m)) (scan 0))" % quota]:
        same_results(ref.run(plan, t), b200.run(plan, t), ordered=False)
