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
