"""STRING / BINARY columns on the CUDA path against the oracle: the reference's own STRING test vectors
(tests/cases.py GOLDEN_STRINGS, pinned to the oracle on the CPU by test_oracle_golden.py) and randomised plans --
scan, filter, comparisons with literals and between columns, group-by / sort / hash join on STRING keys, MIN / MAX /
FIRST / LAST / COUNT of STRING values. Cells are bytes: bit-exact, in order wherever the reference's order is defined."""
import numpy as np
import pytest

from cases import GOLDEN_STRINGS, check_result, same_results
from supersonic_b200 import ssplan as sp

pytestmark = pytest.mark.gpu

WORDS = [b"", b"a", b"ab", b"ab\x00", b"abc", b"b", b"ba", b"zebra", b"0123456789abcdef", b"0123456789abcdefg",
         b"0123456789abcdeg", b"the quick brown fox jumps over the lazy dog", b"the quick brown fox jumps over the lazy cat",
         b"\xff\xfe", b"\x00", b"\x00\x00", b"Z", b"z" * 70]


def string_column(rng, name, n, nullable, vocabulary=None, dtype=sp.STRING):
    if vocabulary is None:   # random cells of random length over a small alphabet: many shared prefixes
        lens = rng.integers(0, 20, n)
        cells = [bytes(rng.integers(97, 100, int(l), dtype=np.uint8)) for l in lens]
    else:
        cells = [vocabulary[i] for i in rng.integers(0, len(vocabulary), n)]
    nulls = (rng.random(n) < 0.15) if nullable else None
    return sp.Column(name, dtype, cells, is_null=nulls)


def string_table(rng, n):
    return [string_column(rng, "s", n, False, WORDS), string_column(rng, "ns", n, True, WORDS),
            string_column(rng, "r", n, False), string_column(rng, "nr", n, True),
            string_column(rng, "b", n, True, WORDS, sp.BINARY),
            sp.Column("k", sp.INT32, rng.integers(0, 7, n).astype(np.int32)),
            sp.Column("v", sp.INT64, rng.integers(-1000, 1000, n)),
            sp.Column("d", sp.DOUBLE, rng.integers(-64, 64, n) / 4.0, is_null=rng.random(n) < 0.1)]


@pytest.mark.parametrize("case", GOLDEN_STRINGS, ids=[c[0] for c in GOLDEN_STRINGS])
@pytest.mark.parametrize("next_rows", [0, 2])
def test_reference_string_vectors(b200, case, next_rows):
    _, plan, tables, expected, ordered = case
    check_result(b200.run(plan, tables, next_max_rows=next_rows), expected, ordered)


@pytest.mark.parametrize("n", [1, 33, 1000, 20011])
def test_scan_filter_compute(ref, b200, n):
    rng = np.random.default_rng(n)
    t = [string_table(rng, n)]
    plans = [
        "(scan 0)",
        "(project (named r nr v) (scan 0))",
        '(filter (equal (col s) (str "abc")) (all) (scan 0))',
        '(filter (less (col nr) (str "b")) (named nr r v) (scan 0))',
        '(filter (greater_or_equal (col ns) (str "0123456789abcdefg")) (named ns k) (scan 0))',
        "(filter (less (col s) (col ns)) (named s ns d) (scan 0))",
        "(filter (equal (col r) (col nr)) (all) (scan 0))",
        '(compute (compound (as a (not_equal (col s) (col r))) (as b (less_or_equal (col nr) (str "bb"))) (as c (is_null (col ns))) (col s) (col nr)) (scan 0))',
        '(compute (compound (as e (equal (col b) (bin "ab"))) (col b)) (filter (greater (col v) (i64 0)) (all) (scan 0)))',
        '(filter (and (greater (col s) (str "a")) (less (col v) (i64 500))) (named s v) (compute (compound (col s) (col v) (as w (plus (col v) (i64 1)))) (scan 0)))',
        '(compute (as x (if (equal (col s) (str "zebra")) (col v) (i64 -1))) (scan 0))',
        '(filter (equal (col s) (str "not in the table")) (all) (scan 0))',
        '(compute (as x (in (col k) (i32 1) (i32 3))) (filter (not_equal (col ns) (str "")) (all) (scan 0)))',
    ]
    for plan in plans:
        for next_rows in (0, 1000):
            same_results(ref.run(plan, t, next_max_rows=next_rows), b200.run(plan, t, next_max_rows=next_rows))


@pytest.mark.parametrize("n", [1, 500, 30011])
def test_group_sort_on_string_keys(ref, b200, n):
    rng = np.random.default_rng(100 + n)
    t = [string_table(rng, n)]
    grouped = [
        ('(group (named s) (aggs (SUM v sv) (COUNT "" c) (MIN d md)) (scan 0))', [0]),
        ('(group (named ns) (aggs (SUM v sv) (COUNT ns c) (MAX r mr) (MIN nr mn)) (scan 0))', [0]),
        ('(group (named nr k) (aggs (COUNT "" c) (MAX s ms)) (scan 0))', [0, 1]),
        ('(group (named k) (aggs (MIN ns a) (MAX ns z) (COUNT ns c) (COUNT b cb)) (scan 0))', [0]),
        ('(group (named r) (aggs (COUNT "" c)) (filter (less (col s) (str "b")) (all) (scan 0)))', [0]),
        ('(group (named b ns) (aggs (SUM v sv)) (scan 0))', [0, 1]),
    ]
    for plan, keys in grouped:
        same_results(ref.run(plan, t), b200.run(plan, t), ordered=False, sort_cols=keys)
    # the reference has no MIN / MAX / FIRST / LAST for BINARY (ERROR_INVALID_ARGUMENT_TYPE at bind time)
    for fn in ("MIN", "MAX", "FIRST", "LAST", "SUM"):
        same_results(ref.run("(group (named k) (aggs (%s b o)) (scan 0))" % fn, t), b200.run("(group (named k) (aggs (%s b o)) (scan 0))" % fn, t))
    # FIRST / LAST depend on the input order only
    same_results(ref.run("(group (named k) (aggs (FIRST nr f) (LAST nr l) (FIRST s fs)) (scan 0))", t),
                 b200.run("(group (named k) (aggs (FIRST nr f) (LAST nr l) (FIRST s fs)) (scan 0))", t), ordered=False, sort_cols=[0])
    same_results(ref.run("(scalar_agg (aggs (MIN r a) (MAX nr z) (COUNT nr c) (COUNT b cb)) (scan 0))", t),
                 b200.run("(scalar_agg (aggs (MIN r a) (MAX nr z) (COUNT nr c) (COUNT b cb)) (scan 0))", t))
    # total orders (the reference's sort is not stable): a unique last key
    t[0].append(sp.Column("id", sp.INT64, np.arange(n)))
    for plan in ["(sort (order (s ASC) (id ASC)) (all) (scan 0))",
                 "(sort (order (nr DESC) (id DESC)) (named nr id r) (scan 0))",
                 "(sort (order (ns ASC) (r DESC) (id ASC)) (named id ns r) (scan 0))",
                 "(sort (order (b DESC) (id ASC)) (named b id) (scan 0))",
                 "(extended_sort (order (r ASC) (id ASC)) 10 (named r id) (scan 0))"]:
        same_results(ref.run(plan, t, next_max_rows=4096), b200.run(plan, t, next_max_rows=4096))


@pytest.mark.parametrize("n", [1, 700, 20000])
def test_hash_join_on_string_keys(ref, b200, n):
    rng = np.random.default_rng(200 + n)
    probe = [string_column(rng, "fk", n, True), string_column(rng, "w", n, False, WORDS), sp.Column("lv", sp.INT64, np.arange(n))]
    m = max(1, n // 3)
    # NOT_UNIQUE build side: random cells (duplicates are likely); UNIQUE build side: distinct cells
    build = [string_column(rng, "pk", m, True), sp.Column("bv", sp.INT64, np.arange(m) * 10), string_column(rng, "bs", m, True, WORDS)]
    distinct = sorted(set(bytes(rng.integers(97, 100, int(l), dtype=np.uint8)) for l in rng.integers(0, 9, m)))
    uniq = [sp.Column("pk", sp.STRING, distinct), sp.Column("bv", sp.INT64, np.arange(len(distinct))),
            sp.Column("bs", sp.STRING, [d[::-1] for d in distinct])]
    for jt in ("INNER", "LEFT_OUTER"):
        plan = "(hash_join %s (named fk) (named pk) (multi (0 (all)) (1 (named bv bs))) NOT_UNIQUE (scan 0) (scan 1))" % jt
        same_results(ref.run(plan, [probe, build], next_max_rows=4096), b200.run(plan, [probe, build], next_max_rows=4096))
        plan = "(hash_join %s (named fk) (named pk) (multi (0 (named fk lv)) (1 (all))) UNIQUE (scan 0) (scan 1))" % jt
        same_results(ref.run(plan, [probe, uniq], next_max_rows=4096), b200.run(plan, [probe, uniq], next_max_rows=4096))
    # two-column key: STRING + INT32
    probe2 = probe + [sp.Column("k", sp.INT32, rng.integers(0, 3, n).astype(np.int32))]
    build2 = build + [sp.Column("k2", sp.INT32, rng.integers(0, 3, m).astype(np.int32))]
    plan = "(hash_join INNER (named fk k) (named pk k2) (multi (0 (named lv fk)) (1 (named bv pk))) NOT_UNIQUE (scan 0) (scan 1))"
    same_results(ref.run(plan, [probe2, build2], next_max_rows=4096), b200.run(plan, [probe2, build2], next_max_rows=4096))
    # a join whose result feeds a group-by on a STRING column of the build side
    plan = ("(group (named bs) (aggs (COUNT \"\" c) (SUM lv s)) (hash_join INNER (named fk) (named pk) "
            "(multi (0 (named lv)) (1 (named bs))) NOT_UNIQUE (scan 0) (scan 1)))")
    same_results(ref.run(plan, [probe, build]), b200.run(plan, [probe, build]), ordered=False, sort_cols=[0])


def test_merge_union_all_and_clusters_with_strings(ref, b200):
    rng = np.random.default_rng(5)
    parts = []
    for i in range(3):
        n = 50 + 25 * i
        cells = sorted(bytes(rng.integers(97, 100, int(l), dtype=np.uint8)) for l in rng.integers(0, 6, n))
        parts.append([sp.Column("s", sp.STRING, cells), sp.Column("src", sp.INT32, np.full(n, i, dtype=np.int32)),
                      sp.Column("id", sp.INT64, np.arange(n))])
    # equal keys of different inputs come out in no defined order in the reference: order by all columns
    plan = "(merge_union_all (order (s ASC) (src ASC) (id ASC)) (scan 0) (scan 1) (scan 2))"
    same_results(ref.run(plan, parts), b200.run(plan, parts))
    plan = '(aggregate_clusters (named s) (aggs (COUNT "" c) (MIN id lo) (MAX src hi)) (merge_union_all (order (s ASC) (src ASC) (id ASC)) (scan 0) (scan 1) (scan 2)))'
    same_results(ref.run(plan, parts), b200.run(plan, parts))


def test_computed_strings_are_refused(b200):
    """Expressions that compute STRING values are outside this slice: refused at run time with ERROR_NOT_IMPLEMENTED,
    never answered differently."""
    t = [[sp.Column("s", sp.STRING, ["a", "b"]), sp.Column("c", sp.BOOL, [True, False])]]
    got = b200.run('(compute (as x (if (col c) (col s) (str "z"))) (scan 0))', t)
    assert got.code == sp.ERROR_NOT_IMPLEMENTED
