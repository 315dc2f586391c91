"""Bind-time parity probe (CPU only): every operator with constants of every type in every operand position, against
the oracle. Prints the expressions whose result code / name / type / nullability differ. Known: constant folding
(DESIGN.md section 6)."""
import sys, itertools; import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from supersonic_b200 import ssplan as sp
import __graft_entry__ as g
from test_binding import COLS, _key, NAMES, BINARY, UNARY, TYPES
ref = sp.PlanLib(g.REF_LIB); b200 = sp.PlanLib(g.PLAN_LIB)
CONSTS = ["(i32 5)", "(i64 -7)", "(u32 3)", "(u64 9)", "(f32 1.5)", "(f64 -2.5)", "(bool true)", "(date 3)", "(datetime 4)",
          "(null INT32)", "(null INT64)", "(null UINT32)", "(null UINT64)", "(null FLOAT)", "(null DOUBLE)", "(null BOOL)", "(null DATE)", "(null DATETIME)"]
exprs = []
for op in BINARY:
    for c in CONSTS:
        for x in ["i32", "ni64", "u32", "f32", "nf64", "b", "d", "ndt"]:
            exprs.append("(%s (col %s) %s)" % (op, x, c))
            exprs.append("(%s %s (col %s))" % (op, c, x))
    for c1, c2 in itertools.product(CONSTS[:9:2] + CONSTS[9::3], repeat=2):
        exprs.append("(%s %s %s)" % (op, c1, c2))
for u in UNARY:
    for c in CONSTS: exprs.append("(%s %s)" % (u, c))
for t in TYPES:
    for c in CONSTS: exprs.append("(cast %s %s)" % (t, c))
for c1, c2 in itertools.product(CONSTS, repeat=2):
    exprs.append("(if (col b) %s %s)" % (c1, c2))
    exprs.append("(if_null %s %s)" % (c1, c2))
for c in CONSTS:
    exprs.append("(if %s (col i32) (col i64))" % c)
    exprs.append("(nulling_if %s (col i32) (col i64))" % c)
    exprs.append("(in (col i64) %s (col i32))" % c)
    exprs.append("(case (col i32) %s (i32 1) (col ni64))" % c)
    exprs.append("(case %s (col i64) (i32 1) (col ni64))" % c)
print(len(exprs), "expressions")
bad = 0
for e in exprs:
    plan = "(compute %s (scan 0))" % e
    a = ref.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY); b = b200.run(plan, [COLS], flags=sp.SSPLAN_BIND_ONLY)
    if _key(a) != _key(b):
        bad += 1
        if bad <= 25: print("MISMATCH", e, "\n   ref:", _key(a), a.error[:140], "\n   b200:", _key(b), b.error[:140])
print(bad, "mismatches")
