"""Bind-time parity of the supersonic.h mirror with the reference: for every (operator, type
pair) the result column name, type, nullability or the error code must match what the
unmodified reference produces (expression/templated/bound_expression_factory.cc:44-123,
expression/core/comparison_bound_expressions.cc:504-636, expression/vector/expression_traits.h
name formats). Runs without a GPU: SSPLAN_BIND_ONLY stops after Operation::CreateCursor."""
import itertools

import numpy as np
import pytest

from supersonic_b200 import ssplan as sp

T = {"i32": sp.INT32, "i64": sp.INT64, "u32": sp.UINT32, "u64": sp.UINT64, "f32": sp.FLOAT,
     "f64": sp.DOUBLE, "b": sp.BOOL, "d": sp.DATE, "dt": sp.DATETIME}
COLS = [sp.Column(n, t, np.ones(4)) for n, t in T.items()]
COLS += [sp.Column("n" + n, t, np.ones(4), is_null=[0, 1, 0, 0]) for n, t in T.items()]
NAMES = [c.name for c in COLS]
BINARY = ("plus minus multiply divide_signaling divide_nulling divide_quiet cpp_divide_signaling "
          "cpp_divide_nulling modulus_signaling modulus_nulling equal not_equal less less_or_equal "
          "greater greater_or_equal and or and_not xor bitwise_and bitwise_or bitwise_xor "
          "bitwise_and_not shift_left shift_right if_null").split()
UNARY = "negate not is_null bitwise_not is_odd is_even".split()
TYPES = ["INT32", "INT64", "UINT32", "UINT64", "FLOAT", "DOUBLE", "BOOL", "DATE", "DATETIME"]


def _key(r):
    return (r.code, r.names, r.dtypes, r.nullable) if r.code == 0 else (r.code,)


def _check(ref, b200, exprs):
    bad = []
    for e in exprs:
        plan = "(compute %s (scan 0))" % e
        a = ref.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
        b = b200.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
        if _key(a) != _key(b):
            bad.append((e, _key(a), _key(b)))
    assert not bad, bad[:10]


@pytest.mark.parametrize("op", BINARY)
def test_binary_binding(ref, b200, op):
    _check(ref, b200, ["(%s (col %s) (col %s))" % (op, x, y) for x, y in itertools.product(NAMES, NAMES)])


def test_unary_and_cast_binding(ref, b200):
    exprs = ["(%s (col %s))" % (u, x) for u in UNARY for x in NAMES]
    exprs += ["(cast %s (col %s))" % (t, x) for t in TYPES for x in NAMES]
    _check(ref, b200, exprs)


def test_if_binding(ref, b200):
    exprs = []
    for x, y in itertools.product(NAMES, NAMES):
        exprs.append("(if (col nb) (col %s) (col %s))" % (x, y))
        exprs.append("(nulling_if (col nb) (col %s) (col %s))" % (x, y))
    exprs += ["(if (col i32) (col i32) (col i32))", "(if (col b) (col i32) (col ni64))"]
    _check(ref, b200, exprs)


def test_constants_aliases_compounds(ref, b200):
    _check(ref, b200, [
        "(plus (i32 5) (i32 6))", "(plus (col i64) (i32 6))", "(plus (col i32) (null INT32))",
        "(less (col i64) (i64 5))", "(multiply (col f64) (i32 2))", "(as foo (plus (col i32) (i32 6)))",
        "(compound (col i32) (as x (col i64)) (plus (col i32) (col i32)))", "(at 2)", "(col nope)", "(at 99)",
        "(plus (plus (col i32) (col i64)) (multiply (col f32) (col u32)))",
        "(and (less (col i32) (i32 3)) (is_null (col ni64)))", "(negate (i32 5))", "(cast INT64 (i32 5))",
        "(is_null (null INT32))", "(if_null (null INT32) (col i32))", "(i32 5)", "(u64 7)", "(f32 1.5)",
        "(f64 2.5)", "(bool true)", "(date 3)", "(datetime 4)", "(null DOUBLE)",
        "(compound (col i32) (col i32))",
    ])


def test_case_and_in_binding(ref, b200):
    exprs = []
    for sw, w in [("i32", "i32"), ("ni32", "i64"), ("i64", "u32"), ("f64", "i32"), ("b", "b"), ("i32", "b"), ("d", "d")]:
        for e, t in [("i64", "ni64"), ("f64", "i32"), ("i32", "i32"), ("b", "nb"), ("i32", "b")]:
            exprs.append("(case (col %s) (col %s) (col %s) (col %s))" % (sw, e, w, t))
            exprs.append("(case (col %s) (col %s) (col %s) (col %s) (col n%s) (col %s))" % (sw, e, w, t, w, e))
    exprs += ["(case (col i32))", "(case (col i32) (col i64) (col i32))", "(case (col i32) (null INT64) (col ni32) (col i64))",
              "(case (col i32) (col i64) (i32 1) (col ni64) (i32 2) (i64 7))"]
    for n, h in itertools.product(["i32", "ni32", "i64", "u64", "f32", "f64", "b", "d"], ["i32", "ni64", "u32", "f64", "nb", "dt"]):
        exprs.append("(in (col %s) (col %s))" % (n, h))
        exprs.append("(in (col %s) (col %s) (col %s))" % (n, h, n))
    exprs += ["(in (col i32) (i32 1) (null INT32))", "(in (col f64) (i32 1) (f64 2.5))", "(in (col i64) (i32 10) (i64 50) (col ni64))"]
    _check(ref, b200, exprs)


def test_operation_schemas(ref, b200):
    """Result schemas of the operators themselves (filter.cc:79-87, aggregator.cc:63-152,
    hash_join.h:37-38, sort.h)."""
    plans = [
        "(filter (less (col i64) (i64 1)) (named i32 f64) (scan 0))",
        "(filter (col i32) (all) (scan 0))",
        "(filter (col nb) (all p_) (scan 0))",
        "(filter (less (col i64) (i64 1)) (named nope) (scan 0))",
        "(project (rename (i32 x) (f64 y)) (scan 0))",
        "(project (at 0 0) (scan 0))",
        "(group (named i32) (aggs (SUM f64 s) (COUNT \"\" c) (MIN ni64 m) (MAX u32 x)) (scan 0))",
        "(group (named ni32 b) (aggs (SUM i32 s INT64) (COUNT ni64 c)) (scan 0))",
        "(group (named i32) (aggs (SUM b s)) (scan 0))",
        "(group (named i32) (aggs (SUM nope s)) (scan 0))",
        "(group (named i32) (aggs (SUM f64 i32)) (scan 0))",
        "(group (named i32) (aggs (COUNT i32 w2 DOUBLE)) (scan 0))",
        "(group (named i32) (aggs (SUM d w2)) (scan 0))",
        "(group (named i32) (aggs (MAX b w2) (SUM u64 w3 INT32) (MIN i32 v INT64)) (scan 0))",
        "(scalar_agg (aggs (SUM f64 s) (COUNT \"\" c)) (scan 0))",
        # round 2: DISTINCT aggregates, the spilling aggregation, Limit, Coalesce, ParseString over literals
        "(group (named i32) (aggs (distinct COUNT ni64 c) (distinct SUM f64 s) (distinct MIN u32 m) (SUM i64 p)) (scan 0))",
        "(group (named i32) (aggs (distinct SUM b s)) (scan 0))",
        "(group (named i32) (aggs (distinct COUNT nope c)) (scan 0))",
        "(scalar_agg (aggs (distinct COUNT i32 c) (distinct SUM ni32 s)) (scan 0))",
        "(hybrid_group 1000 (named ni32 b) (aggs (distinct COUNT f64 c) (MAX u64 x)) (scan 0))",
        "(limit 1 2 (scan 0))",
        "(limit 0 0 (filter (col b) (named i32) (scan 0)))",
        "(coalesce (project (named i32) (scan 0)) (project (rename (i32 again) (nf64 g)) (scan 0)))",
        "(coalesce (project (named i32) (scan 0)) (project (named i32 f64) (scan 0)))",
        "(compute (compound (as d (parse_string_nulling DATE (str \"2001/02/03\"))) (as x (parse_string_quiet UINT64 (str \"7\"))) "
        "(as n (parse_string_nulling FLOAT (str \"x\"))) (col i32)) (scan 0))",
        "(compute (parse_string_nulling INT32 (i32 5)) (scan 0))",
        "(hash_join INNER (named i64) (named i64) (multi (0 (named i32)) (1 (rename (f64 rf)))) UNIQUE (scan 0) (scan 0))",
        "(hash_join LEFT_OUTER (named i64) (named i64) (multi (0 (named i32)) (1 (rename (f64 rf) (ni32 rn)))) NOT_UNIQUE (scan 0) (scan 0))",
        "(hash_join INNER (named i64) (named i32) (multi (0 (named i32)) (1 (rename (f64 rf)))) UNIQUE (scan 0) (scan 0))",
        "(hash_join INNER (named i64) (named i64) (multi (0 (named i32)) (1 (named i32))) UNIQUE (scan 0) (scan 0))",
        "(sort (order (i64 ASC) (f64 DESC)) (named i32 f64) (scan 0))",
        "(sort (order (nope ASC)) (all) (scan 0))",
        "(compute (plus (col e) (i64 1)) (compute (as e (multiply (col i64) (col i64))) (scan 0)))",
        "(filter (less (col e) (i64 1)) (named e) (compute (compound (as e (multiply (col i64) (col i64))) (col i32)) (scan 0)))",
    ]
    bad = []
    for plan in plans:
        a = ref.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
        b = b200.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
        if _key(a) != _key(b):
            bad.append((plan, _key(a), a.error, _key(b), b.error))
    assert not bad, bad


MORE_PLANS = [
    '(scan_selection 0 (ids 3 1 1 0))',
    '(scan_selection 0 (ids))',
    '(filter (col b) (named i32) (scan_selection 0 (ids 2 2)))',
    '(sort (order (i64 ASC)) (all) (scan_selection 0 (ids 0 3)))',
    '(extended_sort (order (i64 ASC) (f64 DESC)) 3 (named i32 f64) (scan 0))',
    '(extended_sort (order (i64 ASC)) none (all) (scan 0))',
    '(extended_sort (order (i64 ASC) (i64 DESC)) none (all) (scan 0))',
    '(extended_sort (order) 0 (all) (scan 0))',
    '(extended_sort (order (ni64 DESC)) 2 (named nope) (scan 0))',
    '(extended_sort (order (ni64 DESC)) 2 (rename (i32 a) (i64 a)) (scan 0))',
    '(extended_sort (order (c DESC)) 1 (all) (group (named i32) (aggs (COUNT "" c)) (scan 0)))',
    '(hash_join INNER (named i64) (named i64) (multi (0 (all)) (1 (all))) UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named i64) (named i64) (multi (0 (all l.)) (1 (all r.))) UNIQUE (scan 0) (scan 0))',
    '(hash_join LEFT_OUTER (named i64) (named i64) (multi (0 (all l.)) (1 (all r.))) UNIQUE (scan 0) (scan 0))',
    '(hash_join LEFT_OUTER (named i64) (named i64) (multi (0 (named i32 ni32)) (1 (rename (i64 k) (b rb) (d rd) (dt rdt) (f32 rf) (u64 ru)))) NOT_UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named i64) (named i64) (multi (0 (at 0 1)) (1 (at 2 3))) UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named i64) (named i64) (multi (0 (at 0)) (1 (at 0))) UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named i64) (named i64) (multi (0 (at 99)) (1 (rename (i32 r)))) UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named i64) (named i64) (multi (0 (named i32)) (1 (at 99))) UNIQUE (scan 0) (scan 0))',
    '(hash_join RIGHT_OUTER (named i64) (named i64) (multi (0 (named i32)) (1 (rename (i32 r)))) UNIQUE (scan 0) (scan 0))',
    '(hash_join FULL_OUTER (named i64) (named i64) (multi (0 (named i32)) (1 (rename (i32 r)))) UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named u64) (named i64) (multi (0 (named i32)) (1 (rename (i32 r)))) UNIQUE (scan 0) (scan 0))',
    '(group (named ni32) (aggs (SUM ni64 s) (MIN nf64 m) (COUNT nb c) (FIRST nd f) (LAST ndt l)) (scan 0))',
    '(group (named ni32 nb nd) (aggs (COUNT "" c)) (scan 0))',
    '(group (all) (aggs (COUNT "" c)) (scan 0))',
    '(group (at 0 1) (aggs (SUM f64 s)) (scan 0))',
    '(group (rename (i32 k)) (aggs (SUM f64 s)) (scan 0))',
    '(group (rename (i32 s)) (aggs (SUM f64 s)) (scan 0))',
    '(group (named i32) (aggs (SUM f64 i64)) (scan 0))',
    '(group (named i32) (aggs (MIN f32 s) (MAX nf32 t) (SUM f32 u) (SUM f32 v DOUBLE)) (scan 0))',
    '(group (named i32) (aggs (SUM u32 s) (SUM u64 t) (SUM i32 u) (SUM nu32 v UINT64)) (scan 0))',
    '(group (named i32) (aggs (MIN u32 s DOUBLE) (MAX i32 t FLOAT) (MIN f64 u FLOAT) (MAX f32 v INT32)) (scan 0))',
    '(group (named i32) (aggs (FIRST i32 s INT64) (LAST f32 t DOUBLE)) (scan 0))',
    '(group (named i32) (aggs (COUNT i32 c UINT32) (COUNT "" d UINT64) (COUNT nf64 e INT32)) (scan 0))',
    '(group (named i32) (aggs (SUM b s INT32)) (scan 0))',
    '(group (named i32) (aggs (MIN dt s) (MAX d t) (MIN b u)) (scan 0))',
    '(group (named i32) (aggs (SUM d s INT32) (SUM dt t INT64)) (scan 0))',
    '(group (named i32) (aggs (MIN d s INT32) (MIN dt t INT64) (MIN i32 u DATE) (MIN i64 v DATETIME)) (scan 0))',
    '(scalar_agg (aggs (SUM ni64 s) (MIN nf64 m) (COUNT nb c) (COUNT "" n) (FIRST b f) (LAST d l)) (scan 0))',
    '(scalar_agg (aggs (SUM f64 s) (SUM f64 s)) (scan 0))',
    '(scalar_agg (aggs (SUM nope s)) (scan 0))',
    '(sort (order (i64 ASC)) (rename (i64 k) (f64 v)) (scan 0))',
    '(sort (order (i64 ASC)) (at 0 0) (scan 0))',
    '(sort (order (i64 ASC)) (named i32 i32) (scan 0))',
    '(sort (order (ni64 DESC) (nb ASC) (nd DESC)) (all s.) (scan 0))',
    '(project (named nope) (scan 0))',
    '(project (at 99) (scan 0))',
    '(project (at 0) (project (at 1) (scan 0)))',
    '(project (rename (i32 a) (i32 b)) (scan 0))',
    '(filter (col nb) (rename (i32 a) (i64 a)) (scan 0))',
    '(filter (and (col b) (col nb)) (at 0 1 2) (scan 0))',
    '(filter (is_null (col ni32)) (named ni32) (scan 0))',
    '(filter (col b) (all) (filter (col nb) (named b i32) (scan 0)))',
    '(compute (compound (as a (col i32)) (as b (plus (col i32) (col ni32)))) (filter (col b) (all) (scan 0)))',
    '(compute (plus (col a) (col b)) (compute (compound (as a (col i32)) (as b (col nf64))) (scan 0)))',
    '(compute (col nope) (compute (as a (col i32)) (scan 0)))',
    '(compute (col i32) (compute (as a (col i32)) (scan 0)))',
    '(group (named k) (aggs (SUM v s)) (hash_join INNER (named i64) (named i64) (multi (0 (rename (i32 k))) (1 (rename (f64 v)))) UNIQUE (scan 0) (scan 0)))',
    '(scalar_agg (aggs (COUNT "" c)) (hash_join LEFT_OUTER (named i64) (named i64) (multi (0 (named i32)) (1 (rename (f64 v)))) NOT_UNIQUE (scan 0) (scan 0)))',
    '(sort (order (v DESC)) (all) (hash_join LEFT_OUTER (named i64) (named i64) (multi (0 (named i32)) (1 (rename (f64 v)))) NOT_UNIQUE (scan 0) (scan 0)))',
    '(filter (is_null (col v)) (all) (hash_join LEFT_OUTER (named i64) (named i64) (multi (0 (named i32)) (1 (rename (f64 v)))) NOT_UNIQUE (scan 0) (scan 0)))',
    '(group (named i32) (aggs (SUM s t)) (group (named i32 i64) (aggs (SUM f64 s)) (scan 0)))',
    '(group (named i32) (aggs (SUM c t)) (group (named i32 i64) (aggs (COUNT "" c)) (scan 0)))',
    '(sort (order (c DESC) (i32 ASC)) (all) (group (named i32) (aggs (COUNT "" c)) (scan 0)))',
    '(compute (divide_signaling (col s) (col c)) (group (named i32) (aggs (SUM f64 s) (COUNT "" c)) (scan 0)))',
    '(compute (divide_nulling (col s) (col c)) (scalar_agg (aggs (SUM i64 s) (COUNT "" c)) (scan 0)))',
    '(bound_compute (compound (as a (col i32)) (as a (col i64))) (bound_scan 0))',
    '(bound_filter (col i32) (all) (bound_scan 0))',
    '(bound_filter (col nb) (named nope) (bound_scan 0))',
    '(bound_group (named i32) (aggs (SUM b s)) (bound_scan 0))',
    '(bound_group (named i32) (aggs (SUM f64 i32)) (bound_scan 0))',
    '(bound_scalar_agg (aggs (SUM nope s)) (bound_scan 0))',
    '(bound_sort (order (i64 ASC) (i64 DESC)) (all) (bound_scan 0))',
    '(bound_project (rename (i32 x) (i64 x)) (bound_scan 0))',
    '(bound_sort (order (c DESC)) (all) (bound_group (named i32) (aggs (COUNT "" c)) (bound_filter (col b) (all) (bound_compute (compound (col i32) (col b)) (bound_scan 0)))))',
    # any join type binds (hash_join.cc:713-726 refuses RIGHT / FULL at the first lookup, not at bind time)
    '(hash_join RIGHT_OUTER (named i64) (named i64) (multi (0 (named i32)) (1 (rename (i32 r)))) UNIQUE (scan 0) (scan 0))',
    '(hash_join FULL_OUTER (named i64) (named i64) (multi (0 (named i32)) (1 (rename (ni32 r)))) NOT_UNIQUE (scan 0) (scan 0))',
    '(group (named i32) (aggs (FIRST f64 s) (LAST ni64 m)) (scan 0))',
    '(group (named i32) (aggs (FIRST b s) (LAST d m) (FIRST dt x)) (scan 0))',
    '(group (named i32) (aggs (FIRST f64 s INT64)) (scan 0))',
    '(group (named i32 i32) (aggs (SUM f64 s)) (scan 0))',
    '(group (named i32) (aggs (SUM f64 s) (SUM f64 s)) (scan 0))',
    '(group (named f64) (aggs (MIN f32 s) (MAX f32 t DOUBLE)) (scan 0))',
    '(group (named b) (aggs (COUNT b c)) (scan 0))',
    '(group (named d dt) (aggs (MIN d a) (MAX dt b2)) (scan 0))',
    '(group (named) (aggs (SUM f64 s)) (scan 0))',
    '(group (named i32) (aggs) (scan 0))',
    '(group (named i32) (aggs (SUM i32 s UINT64)) (scan 0))',
    '(group (named i32) (aggs (SUM u32 s INT32)) (scan 0))',
    '(group (named i32) (aggs (SUM i64 s DOUBLE)) (scan 0))',
    '(group (named i32) (aggs (SUM f64 s INT64)) (scan 0))',
    '(group (named i32) (aggs (MIN i64 s INT32)) (scan 0))',
    '(group (named i32) (aggs (MIN i32 s DOUBLE)) (scan 0))',
    '(group (named i32) (aggs (COUNT "" c INT32)) (scan 0))',
    '(group (named i32) (aggs (COUNT "" c INT64)) (scan 0))',
    '(group (named i32) (aggs (COUNT "" c BOOL)) (scan 0))',
    '(scalar_agg (aggs (MIN b s) (MAX d c) (FIRST ni32 f)) (scan 0))',
    '(scalar_agg (aggs) (scan 0))',
    '(scalar_agg (aggs (SUM dt s)) (scan 0))',
    '(hash_join INNER (named i64 i32) (named i64 i32) (multi (0 (named f64)) (1 (rename (f64 rf)))) UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named f64) (named f64) (multi (0 (named i32)) (1 (rename (i32 r)))) UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named i32) (named u32) (multi (0 (named i32)) (1 (rename (i32 r)))) UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named ni64) (named i64) (multi (0 (named i32)) (1 (rename (i32 r)))) UNIQUE (scan 0) (scan 0))',
    '(hash_join LEFT_OUTER (named ni64) (named ni64) (multi (0 (all)) (1 (rename (i32 r) (ni32 rn) (b rb)))) UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named nope) (named i64) (multi (0 (named i32)) (1 (rename (i32 r)))) UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named i64) (named i64) (multi (0 (named nope)) (1 (rename (i32 r)))) UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named d) (named d) (multi (0 (named i32)) (1 (rename (i32 r)))) NOT_UNIQUE (scan 0) (scan 0))',
    '(hash_join INNER (named b) (named b) (multi (0 (named i32)) (1 (rename (i32 r)))) NOT_UNIQUE (scan 0) (scan 0))',
    '(sort (order (i64 ASC) (i64 DESC)) (all) (scan 0))',
    '(sort (order (b ASC) (d DESC) (dt ASC) (f32 DESC) (u64 ASC)) (all) (scan 0))',
    '(sort (order (ni64 ASC)) (named ni64 nf64) (scan 0))',
    '(sort (order (i64 ASC)) (named nope) (scan 0))',
    '(sort (order) (all) (scan 0))',
    '(project (named i32 i32) (scan 0))',
    '(project (rename (i32 x) (i64 x)) (scan 0))',
    '(project (all) (project (named i32 nf64) (scan 0)))',
    '(project (all p.) (scan 0))',
    '(filter (i32 1) (all) (scan 0))',
    '(filter (bool true) (all) (scan 0))',
    '(filter (null BOOL) (all) (scan 0))',
    '(filter (col b) (named i32 i32) (scan 0))',
    '(filter (less (col nope) (i64 1)) (all) (scan 0))',
    '(compute (compound (col i32) (col i32)) (scan 0))',
    '(compute (compound (as a (col i32)) (as a (col i64))) (scan 0))',
    '(compute (col i32) (filter (col b) (named i32 i64) (scan 0)))',
    '(group (named e) (aggs (SUM i64 s)) (compute (compound (as e (plus (col i32) (i32 1))) (col i64)) (scan 0)))',
    '(group (named i32) (aggs (SUM e s)) (filter (col b) (named i32 e) (compute (compound (col i32) (col b) (as e (multiply (col f64) (col f64)))) (scan 0))))',
    '(sort (order (s DESC)) (all) (group (named i32) (aggs (SUM f64 s)) (scan 0)))',
    '(hash_join INNER (named i32) (named k) (multi (0 (named i64)) (1 (named s))) UNIQUE (scan 0) (group (named k) (aggs (SUM f64 s)) (project (rename (i32 k) (f64 f64)) (scan 0))))',
    "(bound_scan 0)', '(bound_compute (col i32) (bound_scan 0))",
    '(bound_filter (less (col i64) (i64 1)) (named i32) (bound_scan 0))',
    '(bound_group (named i32) (aggs (SUM f64 s) (FIRST ni32 f)) (bound_scan 0))',
    '(bound_scalar_agg (aggs (SUM f64 s)) (bound_scan 0))',
    '(bound_sort (order (i64 DESC)) (named i32) (bound_scan 0))',
    '(bound_project (named nope) (bound_scan 0))',
    '(bound_group (named nope) (aggs (SUM f64 s)) (bound_scan 0))',
    '(bound_sort (order (nope DESC)) (all) (bound_scan 0))',
]


def test_more_operation_schemas(ref, b200):
    """FIRST / LAST and result-type overrides of the aggregates, key and payload schemas of the join, duplicate
    and missing sort keys, nested plans and the cursor-level (bound_*) factories."""
    bad = []
    for plan in MORE_PLANS:
        a = ref.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
        b = b200.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
        if _key(a) != _key(b):
            bad.append((plan, _key(a), a.error, _key(b), b.error))
    assert not bad, bad


def test_join_key_schemas_are_checked_at_bind_time(ref, b200):
    """Deliberate divergence: the reference only DCHECKs that the two key schemas agree (hash_join.cc:383-387,
    490, 710) and a release build then compares keys of different types or counts bytewise; the mirror refuses such
    plans with ERROR_ATTRIBUTE_TYPE_MISMATCH. Integer keys of different width or signedness stay legal."""
    for plan in ["(hash_join INNER (named i64 i32) (named i64) (multi (0 (named f64)) (1 (rename (f64 rf)))) UNIQUE (scan 0) (scan 0))",
                 "(hash_join INNER (named f64) (named f32) (multi (0 (named i32)) (1 (rename (i32 r)))) UNIQUE (scan 0) (scan 0))"]:
        assert ref.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY).code == 0
        assert b200.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY).code == 402
    ok = "(hash_join INNER (named i32) (named u32) (multi (0 (named i32)) (1 (rename (i32 r)))) UNIQUE (scan 0) (scan 0))"
    assert ref.run(ok, [COLS], flags=sp.SSPLAN_BIND_ONLY).code == b200.run(ok, [COLS], flags=sp.SSPLAN_BIND_ONLY).code == 0


@pytest.mark.parametrize("agg", ["SUM", "MIN", "MAX", "COUNT", "FIRST", "LAST"])
def test_aggregate_binding_cross_product(ref, b200, agg):
    """Every aggregate over every column type and nullability, with and without a result-type override, grouped and
    scalar (aggregator.cc:63-152): result type, nullability or the error code."""
    bad = []
    for x in NAMES:
        for t in [""] + TYPES:
            for shape in ['(group (named i32) (aggs (%s %s r%s)) (scan 0))', '(scalar_agg (aggs (%s %s r%s)) (scan 0))']:
                plan = shape % (agg, x, (" " + t) if t else "")
                a = ref.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
                b = b200.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
                if _key(a) != _key(b):
                    bad.append((plan, _key(a), _key(b)))
    assert not bad, bad[:10]


def test_operators_over_every_column_type(ref, b200):
    """Filter predicates, sort keys, group keys, join keys (both join types and uniqueness declarations, LEFT_OUTER
    nullability of the build side's columns), projections and compound expressions over every column type."""
    plans = []
    for x in NAMES:
        plans.append('(filter (col %s) (all) (scan 0))' % x)
        plans.append('(filter (is_null (col %s)) (named %s) (scan 0))' % (x, x))
        plans.append('(sort (order (%s ASC)) (named %s) (scan 0))' % (x, x))
        plans.append('(sort (order (%s DESC) (i32 ASC)) (all) (scan 0))' % x)
        plans.append('(extended_sort (order (%s DESC)) 5 (all) (scan 0))' % x)
        plans.append('(group (named %s) (aggs (COUNT "" c)) (scan 0))' % x)
        plans.append('(group (named %s i64) (aggs (SUM f64 s) (MIN %s m)) (scan 0))' % (x, x))
        for jt in ("INNER", "LEFT_OUTER"):
            for u in ("UNIQUE", "NOT_UNIQUE"):
                plans.append('(hash_join %s (named %s) (named %s) (multi (0 (named i32)) (1 (rename (%s r) (nf64 rn) (b rb)))) %s '
                             '(scan 0) (scan 0))' % (jt, x, x, x, u))
        plans.append('(project (rename (%s a) (i32 b)) (scan 0))' % x)
        plans.append('(compute (compound (as a (col %s)) (as b (is_null (col %s)))) (scan 0))' % (x, x))
    bad = []
    for plan in plans:
        a = ref.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
        b = b200.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
        if _key(a) != _key(b):
            bad.append((plan, _key(a), a.error, _key(b), b.error))
    assert not bad, bad[:10]


# ---- the bound factories (BoundNamedAttribute, BoundConst*, BoundPlus, BoundLess, BoundIf, BoundCastTo, BoundAlias,
# BoundCompoundExpression ...: expression/core/*_bound_expressions.h, terminal_bound_expressions.h) assembled bottom-up
# by the plan driver's BuildBoundExpr and wrapped by CreateBoundExpressionTree: same names, types, nullability and
# error codes as the reference's own factories.
def _check_bound(ref, b200, exprs):
    bad = []
    for e in exprs:
        plan = "(bound_bx_compute %s (bound_scan 0))" % e
        a = ref.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
        b = b200.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
        if _key(a) != _key(b):
            bad.append((e, _key(a), _key(b)))
    assert not bad, bad[:10]


@pytest.mark.parametrize("op", BINARY)
def test_bound_factory_binary_binding(ref, b200, op):
    some = ["i32", "i64", "u32", "u64", "f32", "f64", "b", "d", "dt", "ni32", "nf64", "nb"]
    _check_bound(ref, b200, ["(%s (col %s) (col %s))" % (op, x, y) for x, y in itertools.product(some, some)])


def test_bound_factory_other_binding(ref, b200):
    exprs = ["(%s (col %s))" % (u, x) for u in UNARY for x in NAMES]
    exprs += ["(cast %s (col %s))" % (t, x) for t in TYPES for x in NAMES]
    exprs += ["(if (col nb) (col %s) (col %s))" % (x, y) for x, y in itertools.product(NAMES[:9], NAMES[9:])]
    exprs += ["(nulling_if (col b) (col i32) (col ni64))", "(as renamed (plus (col i32) (i32 1)))", "(plus (i64 1) (i32 2))",
              "(compound (col i32) (as e (multiply (col f64) (f64 2))) (less (col i32) (i32 5)))", "(compound (col i32) (col i32))",
              "(plus (col i32) (null INT32))", "(is_null (null DOUBLE))", "(col missing)", "(at 99)", "(at 3)",
              "(if_null (col ni32) (i32 0))", "(equal (col i32) (u64 7))"]
    _check_bound(ref, b200, exprs)
