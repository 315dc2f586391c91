"""Generates the committed golden fixtures of tests/golden/ from the oracle (the unmodified reference,
oracle/_ref/libssref.so): seeded inputs -> the reference's outputs for the plans of the hot path.

    python tests/golden/make_golden.py          (needs oracle/_ref, i.e. /root/reference to build it)

tests/test_golden_fixtures.py then checks the oracle (CPU) and the B200 path (GPU) against these
files, so a change in either shows up even where /root/reference is absent."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from supersonic_b200 import ssplan as sp  # noqa: E402


def tables(seed=2024):
    """The seeded input tables of all fixtures (small enough to commit their outputs)."""
    rng = np.random.default_rng(seed)
    n = 20_000
    t0 = [sp.Column("a", sp.INT64, rng.integers(-2**31, 2**31, n)), sp.Column("b", sp.INT64, rng.integers(-2**31, 2**31, n)),
          sp.Column("c", sp.INT64, rng.integers(-2**62, 2**62, n)), sp.Column("d", sp.INT64, rng.integers(0, 2**20, n)),
          sp.Column("k", sp.INT64, rng.integers(0, 37, n), is_null=rng.random(n) < 0.02),
          sp.Column("v", sp.DOUBLE, rng.integers(0, 1 << 20, n) / 1024.0, is_null=rng.random(n) < 0.1),
          sp.Column("i", sp.INT32, rng.integers(-1000, 1000, n).astype(np.int32))]
    m = 3_000
    t1 = [sp.Column("pk", sp.INT64, rng.permutation(m).astype(np.int64) * 7), sp.Column("payload", sp.INT64, rng.integers(0, 10**9, m)),
          sp.Column("w", sp.DOUBLE, rng.random(m), is_null=rng.random(m) < 0.1)]
    t0.append(sp.Column("fk", sp.INT64, rng.integers(0, m * 7, n), is_null=rng.random(n) < 0.03))
    return [t0, t1]


# name -> (plan, ordered?)
PLANS = {
    "c1_filter_project": ("(filter (less (col d) (i64 524288)) (named e) (compute (compound "
                          "(as e (plus (multiply (col a) (col b)) (col c))) (col d)) (scan 0)))", True),
    "expressions": ("(compute (compound (as p (plus (col a) (col i))) (as q (divide_nulling (col c) (col d))) "
                    "(as r (if (less (col v) (f64 500)) (col a) (col b))) (as s (and (is_null (col k)) (greater (col i) (i32 0)))) "
                    "(as t (cast DOUBLE (col i)))) (scan 0))", True),
    "group_by": ("(group (named k) (aggs (SUM v sv) (MIN a mn) (MAX i mx) (COUNT v cv) (COUNT \"\" n) (FIRST a fa) (LAST i li)) (scan 0))", False),
    "scalar": ("(scalar_agg (aggs (SUM i si) (MAX v mv) (COUNT \"\" n)) (scan 0))", True),
    "q1_shape": ("(group (named k) (aggs (SUM x sx) (COUNT \"\" n)) (compute (compound (col k) (as x (multiply (col i) (plus (col i) (i32 1))))) "
                 "(filter (less_or_equal (col d) (i64 900000)) (all) (scan 0))))", False),
    "join_inner": ("(hash_join INNER (named fk) (named pk) (multi (0 (named fk i)) (1 (named payload w))) UNIQUE (scan 0) (scan 1))", True),
    "join_left_outer": ("(hash_join LEFT_OUTER (named fk) (named pk) (multi (0 (named fk i)) (1 (named payload w))) NOT_UNIQUE (scan 0) (scan 1))", True),
    "sort": ("(sort (order (k ASC) (i DESC) (a ASC)) (named k i a v) (scan 0))", True),
}


def main():
    ref = sp.PlanLib(os.path.join(ROOT, "oracle", "_ref", "libssref.so"))
    tabs = tables()
    for name, (plan, _) in PLANS.items():
        r = ref.run(plan, tabs)
        assert r.code == 0, (name, r.error)
        out = {"names": np.array(r.names), "dtypes": np.array(r.dtypes, dtype=np.int32), "rows": np.int64(r.rows)}
        for j, col in enumerate(r.columns):
            isn = r.nulls[j] if r.nulls[j] is not None else np.zeros(r.rows, dtype=bool)
            col = col.copy()
            if col.dtype == np.bool_:
                col = col.astype(np.uint8)
            col[isn] = 0                      # values under NULL are unspecified: store zeros
            out["col%d" % j] = col
            out["null%d" % j] = np.packbits(isn)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("%-20s %6d rows" % (name, r.rows))


if __name__ == "__main__":
    main()
