"""Shared test material: known-answer vectors taken from the reference's own tests (cited per
case) and randomised plan families. `N` marks a NULL cell."""
import numpy as np

from supersonic_b200 import ssplan as sp

N = None


def col(name, dtype, values):
    """Column from a python list; None entries become NULL."""
    nulls = [v is None for v in values]
    data = [0 if v is None else v for v in values]
    if any(nulls):
        return sp.Column(name, dtype, data, is_null=nulls)
    return sp.Column(name, dtype, data)


def ncol(name, dtype, values):
    """As col() but always NULLABLE."""
    nulls = [v is None for v in values]
    data = [0 if v is None else v for v in values]
    return sp.Column(name, dtype, data, is_null=nulls)


# (id, plan, tables, expected {name: list}, ordered)
GOLDEN = [
    # expression/core/arithmetic_expressions_test.cc:68-99 (Plus INT64)
    ("plus_int64", "(compute (plus (col a) (col b)) (scan 0))",
     [[col("a", sp.INT64, [-1, -2, 2, 13]), col("b", sp.INT64, [1, 2, 2, 1])]],
     {"(a + b)": [0, 0, 4, 14]}, True),
    # :80-92 nullable rows
    ("plus_nullable", "(compute (plus (col a) (col b)) (scan 0))",
     [[ncol("a", sp.INT64, [1, N, 3, N]), ncol("b", sp.INT64, [1, 2, N, N])]],
     {"(a + b)": [2, N, N, N]}, True),
    # :94-99 INT64 + INT32 promotes
    ("plus_promote", "(compute (plus (col a) (col b)) (scan 0))",
     [[col("a", sp.INT64, [5, -7]), col("b", sp.INT32, [3, 7])]],
     {"(a + CAST_INT32_TO_INT64(b))": [8, 0]}, True),
    # :127-144 Minus / Multiply incl. NULL rows
    ("minus", "(compute (minus (col a) (col b)) (scan 0))",
     [[ncol("a", sp.INT32, [3, 10, N]), ncol("b", sp.INT32, [1, 20, 4])]],
     {"(a - b)": [2, -10, N]}, True),
    ("multiply", "(compute (multiply (col a) (col b)) (scan 0))",
     [[ncol("a", sp.INT32, [20, N, -3]), ncol("b", sp.INT32, [20, 2, 4])]],
     {"(a * b)": [400, N, -12]}, True),
    # :146-218 Divide quiet / nulling
    ("divide_quiet", "(compute (divide_quiet (col a) (col b)) (scan 0))",
     [[col("a", sp.DOUBLE, [1.0, 0.0, 6.0]), col("b", sp.DOUBLE, [0.0, 0.0, 4.0])]],
     {"(a /. b)": [float("inf"), float("nan"), 1.5]}, True),
    ("divide_nulling", "(compute (divide_nulling (col a) (col b)) (scan 0))",
     [[col("a", sp.INT32, [1, 0, 6]), col("b", sp.INT32, [0, 0, 4])]],
     {"(CAST_INT32_TO_DOUBLE(a) /. CAST_INT32_TO_DOUBLE(b))": [N, N, 1.5]}, True),
    ("cpp_divide_nulling", "(compute (cpp_divide_nulling (col a) (col b)) (scan 0))",
     [[col("a", sp.INT32, [7, -7, 5]), col("b", sp.INT32, [2, 2, 0])]],
     {"(a / b)": [3, -3, N]}, True),
    ("modulus_nulling", "(compute (modulus_nulling (col a) (col b)) (scan 0))",
     [[col("a", sp.INT32, [7, -7, 5]), col("b", sp.INT32, [3, 3, 0])]],
     {"(a % b)": [1, -1, N]}, True),
    # cursor/core/compute_test.cc:51-79
    ("compute_sum", "(compute (as sum (plus (col col1) (col col3))) (scan 0))",
     [[ncol("col1", sp.INT64, [12, 13, N]), col("col2", sp.INT64, [1, 2, 3]), ncol("col3", sp.INT64, [5, 6, N])]],
     {"sum": [17, 19, N]}, True),
    # elementary_expressions_test.cc: three-valued logic truth tables
    ("and_3vl", "(compute (and (col a) (col b)) (scan 0))",
     [[ncol("a", sp.BOOL, [True, True, True, False, False, False, N, N, N]),
       ncol("b", sp.BOOL, [True, False, N, True, False, N, True, False, N])]],
     {"(a AND b)": [True, False, N, False, False, False, N, False, N]}, True),
    ("or_3vl", "(compute (or (col a) (col b)) (scan 0))",
     [[ncol("a", sp.BOOL, [True, True, True, False, False, False, N, N, N]),
       ncol("b", sp.BOOL, [True, False, N, True, False, N, True, False, N])]],
     {"(a OR b)": [True, True, True, True, False, N, True, N, N]}, True),
    ("not_3vl", "(compute (not (col a)) (scan 0))",
     [[ncol("a", sp.BOOL, [True, False, N])]], {"(NOT a)": [False, True, N]}, True),
    ("if_null_cond", "(compute (if (col c) (col a) (col b)) (scan 0))",
     [[ncol("c", sp.BOOL, [True, False, N]), col("a", sp.INT32, [1, 2, 3]), col("b", sp.INT32, [10, 20, 30])]],
     {"IF c THEN a ELSE b": [1, 20, 30]}, True),
    ("ifnull", "(compute (if_null (col a) (col b)) (scan 0))",
     [[ncol("a", sp.INT32, [1, N, N]), ncol("b", sp.INT32, [5, 6, N])]],
     {"IFNULL(a, b)": [1, 6, N]}, True),
    # expression/core/case_expression_test.cc and comparison IN (comparison_expressions.h:76-88)
    ("case_switch", "(compute (case (col s) (col a) (i32 1) (col b) (i32 2) (i64 7)) (scan 0))",
     [[ncol("s", sp.INT32, [1, 2, 3, N, 5]), col("a", sp.INT64, [10, 20, 30, 40, 50]), ncol("b", sp.INT64, [100, N, 300, 400, 500])]],
     {"CASE(s, a, CONST_INT32, b, CONST_INT32, CONST_INT64)": [100, 7, 30, 40, 50]}, True),
    ("case_null_else", "(compute (case (col s) (null INT64) (col w) (col a)) (scan 0))",
     [[ncol("s", sp.INT32, [1, 2, 3, N, 5]), col("a", sp.INT64, [10, 20, 30, 40, 50]), ncol("w", sp.INT32, [1, N, 3, 4, 9])]],
     {"CASE(s, NULL, w, a)": [10, N, 30, N, N]}, True),
    ("in_consts", "(compute (in (col s) (i32 1) (i32 5)) (scan 0))",
     [[ncol("s", sp.INT32, [1, 2, 3, N, 5])]], {"s IN (CONST_INT32, CONST_INT32)": [True, False, False, N, True]}, True),
    ("in_null_element", "(compute (in (col s) (i32 1) (null INT32)) (scan 0))",
     [[ncol("s", sp.INT32, [1, 2, 3, N, 5])]], {"s IN (CONST_INT32, NULL)": [True, N, N, N, N]}, True),
    ("in_columns", "(compute (in (col s) (col w) (i32 2)) (scan 0))",
     [[ncol("s", sp.INT32, [1, 2, 3, N, 5]), ncol("w", sp.INT32, [1, N, 3, 4, 9])]],
     {"s IN (w, CONST_INT32)": [True, True, True, N, False]}, True),
    # cursor/core/filter_test.cc:151-329
    ("filter_some", "(filter (col p) (named v) (scan 0))",
     [[col("p", sp.BOOL, [True, False, True, False, True]), col("v", sp.INT32, [1, 2, 3, 4, 5])]],
     {"v": [1, 3, 5]}, True),
    ("filter_none", "(filter (col p) (named v) (scan 0))",
     [[col("p", sp.BOOL, [False, False]), col("v", sp.INT32, [1, 2])]], {"v": []}, True),
    ("filter_null_predicate", "(filter (col p) (all) (scan 0))",
     [[ncol("p", sp.BOOL, [True, N, False, True]), ncol("v", sp.INT64, [1, 2, 3, N])]],
     {"p": [True, True], "v": [1, N]}, True),
    ("filter_projected_away", "(filter (less (col k) (i32 3)) (named v) (scan 0))",
     [[col("k", sp.INT32, [1, 5, 2, 7]), col("v", sp.DOUBLE, [0.5, 1.5, 2.5, 3.5])]],
     {"v": [0.5, 2.5]}, True),
    # cursor/core/aggregate_groups_test.cc:102-218, :272-294, :330-353, :379-428
    ("group_sum", "(group (named k) (aggs (SUM v sum) (COUNT v cnt) (COUNT \"\" cnt_all)) (scan 0))",
     [[col("k", sp.INT32, [1, 3, 1, 3, 1]), ncol("v", sp.INT32, [3, -3, 4, -5, N])]],
     {"k": [1, 3], "sum": [7, -8], "cnt": [2, 2], "cnt_all": [3, 2]}, False),
    ("group_all_null_sum", "(group (named k) (aggs (SUM v sum) (MIN v mn) (MAX v mx)) (scan 0))",
     [[col("k", sp.INT32, [1, 1, 2]), ncol("v", sp.INT64, [N, N, 9])]],
     {"k": [1, 2], "sum": [N, 9], "mn": [N, 9], "mx": [N, 9]}, False),
    ("group_null_key", "(group (named k) (aggs (SUM v s)) (scan 0))",
     [[ncol("k", sp.INT32, [1, N, 1, N]), col("v", sp.INT32, [1, 2, 3, 4])]],
     {"k": [N, 1], "s": [6, 4]}, False),
    ("group_two_keys", "(group (named a b) (aggs (SUM v s) (MIN v m) (COUNT \"\" c)) (scan 0))",
     [[col("a", sp.INT32, [1, 1, 2, 1, 2]), col("b", sp.INT64, [7, 8, 7, 7, 7]), col("v", sp.DOUBLE, [1.0, 2.0, 3.0, 4.0, 5.0])]],
     {"a": [1, 1, 2], "b": [7, 8, 7], "s": [5.0, 2.0, 8.0], "m": [1.0, 2.0, 3.0], "c": [2, 1, 2]}, False),
    ("group_sum_int32_to_int64", "(group (named k) (aggs (SUM v s INT64)) (scan 0))",
     [[col("k", sp.INT32, [1, 1]), col("v", sp.INT32, [2000000000, 2000000000])]],
     {"k": [1], "s": [4000000000]}, False),
    ("group_empty", "(group (named k) (aggs (SUM v s)) (scan 0))",
     [[col("k", sp.INT32, []), col("v", sp.INT32, [])]], {"k": [], "s": []}, False),
    # cursor/core/aggregate_scalar_test.cc:53-90
    ("scalar_agg", "(scalar_agg (aggs (SUM v s) (COUNT \"\" c) (MAX v m)) (scan 0))",
     [[col("v", sp.INT64, [5, 7, -2])]], {"s": [10], "c": [3], "m": [7]}, True),
    ("scalar_agg_empty", "(scalar_agg (aggs (SUM v s) (COUNT \"\" c)) (scan 0))",
     [[col("v", sp.INT64, [])]], {"s": [N], "c": [0]}, True),
    # test/guide/primer.cc GroupAggregateTest
    ("primer_group", "(group (named key) (aggs (SUM data data_sums)) (scan 0))",
     [[col("key", sp.INT32, [1, 2, 3, 1, 2, 3, 1, 2]), col("data", sp.DOUBLE, [1.5, 3.0, 3.0, 7.6, 5.5, 2.0, 1.6, 9.5])]],
     {"key": [1, 2, 3], "data_sums": [1.5 + 7.6 + 1.6, 3.0 + 5.5 + 9.5, 3.0 + 2.0]}, False),
    # cursor/core/hash_join_test.cc:140-238, 240-281, 305-319, 355-382
    ("join_inner_unique",
     "(hash_join INNER (named k) (named k2) (multi (0 (all)) (1 (named w))) UNIQUE (scan 0) (scan 1))",
     [[col("k", sp.INT64, [1, 2, 3, 4, 5]), col("v", sp.INT64, [10, 20, 30, 40, 50])],
      [col("k2", sp.INT64, [6, 5, 4, 3, 2, 1]), col("w", sp.INT64, [600, 500, 400, 300, 200, 100])]],
     {"k": [1, 2, 3, 4, 5], "v": [10, 20, 30, 40, 50], "w": [100, 200, 300, 400, 500]}, True),
    ("join_left_outer",
     "(hash_join LEFT_OUTER (named k) (named k2) (multi (0 (all)) (1 (named w))) UNIQUE (scan 0) (scan 1))",
     [[col("k", sp.INT64, [1, 2, 3]), col("v", sp.INT64, [10, 20, 30])],
      [col("k2", sp.INT64, [3, 1]), col("w", sp.INT64, [300, 100])]],
     {"k": [1, 2, 3], "v": [10, 20, 30], "w": [100, N, 300]}, True),
    ("join_duplicates",
     "(hash_join INNER (named k) (named k2) (multi (0 (named k v)) (1 (named w))) NOT_UNIQUE (scan 0) (scan 1))",
     [[col("k", sp.INT32, [2, 3, 2]), col("v", sp.INT32, [1, 2, 3])],
      [col("k2", sp.INT32, [2, 2, 3, 2]), col("w", sp.INT32, [10, 20, 30, 40])]],
     {"k": [2, 2, 2, 3, 2, 2, 2], "v": [1, 1, 1, 2, 3, 3, 3], "w": [10, 20, 40, 30, 10, 20, 40]}, True),
    ("join_two_keys",
     "(hash_join INNER (named a b) (named a2 b2) (multi (0 (named v)) (1 (named w))) UNIQUE (scan 0) (scan 1))",
     [[col("a", sp.INT32, [1, 1, 2]), col("b", sp.INT64, [1, 2, 1]), col("v", sp.INT32, [10, 20, 30])],
      [col("a2", sp.INT32, [2, 1]), col("b2", sp.INT64, [1, 2]), col("w", sp.INT32, [7, 8])]],
     {"v": [20, 30], "w": [8, 7]}, True),
    ("join_null_keys",
     "(hash_join LEFT_OUTER (named k) (named k2) (multi (0 (named v)) (1 (named w))) NOT_UNIQUE (scan 0) (scan 1))",
     [[ncol("k", sp.INT32, [1, N, 2]), col("v", sp.INT32, [10, 20, 30])],
      [ncol("k2", sp.INT32, [N, 1, N]), col("w", sp.INT32, [7, 8, 9])]],
     {"v": [10, 20, 30], "w": [8, N, N]}, True),
    ("join_empty_rhs",
     "(hash_join INNER (named k) (named k2) (multi (0 (named v)) (1 (named w))) UNIQUE (scan 0) (scan 1))",
     [[col("k", sp.INT32, [1, 2]), col("v", sp.INT32, [10, 20])], [col("k2", sp.INT32, []), col("w", sp.INT32, [])]],
     {"v": [], "w": []}, True),
    # cursor/core/sort_test.cc:121-381
    ("sort_asc_nulls_first", "(sort (order (k ASC)) (all) (scan 0))",
     [[ncol("k", sp.INT32, [3, N, 1, 2]), col("v", sp.INT32, [30, 0, 10, 20])]],
     {"k": [N, 1, 2, 3], "v": [0, 10, 20, 30]}, True),
    ("sort_desc_nulls_last", "(sort (order (k DESC)) (all) (scan 0))",
     [[ncol("k", sp.INT32, [3, N, 1, 2]), col("v", sp.INT32, [30, 0, 10, 20])]],
     {"k": [3, 2, 1, N], "v": [30, 20, 10, 0]}, True),
    ("sort_two_keys", "(sort (order (a ASC) (b DESC)) (named b a) (scan 0))",
     [[col("a", sp.INT64, [2, 1, 2, 1]), col("b", sp.DOUBLE, [0.5, -1.0, 7.0, 3.0])]],
     {"b": [3.0, -1.0, 7.0, 0.5], "a": [1, 1, 2, 2]}, True),
]

# Cases added after the last GPU run of round 1: oracle-pinned on the CPU; test_parity_gpu.py runs them last.
GOLDEN_LATE = [
    # ... but they are two hash keys: the row hash (hash of the bit image) tells them apart before == is asked
    ("group_signed_zero_keys", "(group (named x) (aggs (SUM v s) (COUNT \"\" c)) (scan 0))",
     [[col("x", sp.DOUBLE, [0.0, -0.0, 1.0, -0.0, 0.0]), col("v", sp.INT64, [0, 1, 2, 3, 4])]],
     {"x": [0.0, -0.0, 1.0], "s": [4, 4, 2], "c": [2, 2, 1]}, False),
    ("join_signed_zero_keys",
     "(hash_join INNER (named x) (named y) (multi (0 (named v)) (1 (named w))) UNIQUE (scan 0) (scan 1))",
     [[col("x", sp.DOUBLE, [0.0, -0.0, 1.0, -0.0, 0.0]), col("v", sp.INT64, [0, 1, 2, 3, 4])],
      [col("y", sp.DOUBLE, [-0.0, 2.0]), col("w", sp.INT64, [7, 8])]],
     {"v": [1, 3], "w": [7, 7]}, True),
    # scan_view.h:37-46: a cursor over view[selection[i]]; rows may be selected any number of times
    ("scan_selection", "(scan_selection 0 (ids 4 1 1 0))",
     [[ncol("k", sp.INT32, [3, N, 2, 5, 4]), col("v", sp.DOUBLE, [30.0, 10.0, 20.0, 50.0, 40.0])]],
     {"k": [4, N, N, 3], "v": [40.0, 10.0, 10.0, 30.0]}, True),
    ("scan_selection_empty", "(scan_selection 0 (ids))",
     [[ncol("k", sp.INT32, [3, N, 2]), col("v", sp.DOUBLE, [30.0, 10.0, 20.0])]],
     {"k": [], "v": []}, True),
    ("scan_selection_compute", "(compute (as e (plus (col v) (col k))) (scan_selection 0 (ids 2 1 2)))",
     [[ncol("k", sp.INT32, [3, N, 2, 5, 4]), col("v", sp.DOUBLE, [30.0, 10.0, 20.0, 50.0, 40.0])]],
     {"e": [22.0, N, 22.0]}, True),
    ("scan_selection_group", "(group (named k) (aggs (SUM v s)) (scan_selection 0 (ids 4 4 1 1 0)))",
     [[ncol("k", sp.INT32, [3, N, 2, 5, 4]), col("v", sp.DOUBLE, [30.0, 10.0, 20.0, 50.0, 40.0])]],
     {"k": [4, N, 3], "s": [80.0, 20.0, 30.0]}, False),
    # sort.h:103-131, sort.cc:857-1017: sort by attribute names, first `limit` rows (cursor/core/sort_test.cc ExtendedSort cases)
    ("extended_sort_limit", "(extended_sort (order (k DESC)) 3 (all) (scan 0))",
     [[col("k", sp.INT32, [3, 1, 2, 5, 4]), col("v", sp.INT32, [30, 10, 20, 50, 40])]],
     {"k": [5, 4, 3], "v": [50, 40, 30]}, True),
    ("extended_sort_no_limit", "(extended_sort (order (a ASC) (b DESC)) none (named b a) (scan 0))",
     [[col("a", sp.INT64, [2, 1, 2, 1]), col("b", sp.DOUBLE, [0.5, -1.0, 7.0, 3.0])]],
     {"b": [3.0, -1.0, 7.0, 0.5], "a": [1, 1, 2, 2]}, True),
    ("extended_sort_limit_beyond_rows", "(extended_sort (order (k ASC)) 9 (named v) (scan 0))",
     [[ncol("k", sp.INT32, [3, N, 1, 2]), col("v", sp.INT32, [30, 0, 10, 20])]],
     {"v": [0, 10, 20, 30]}, True),
    ("extended_sort_limit_zero", "(extended_sort (order (k ASC)) 0 (all) (scan 0))",
     [[col("k", sp.INT32, [3, 1, 2]), col("v", sp.INT32, [30, 10, 20])]],
     {"k": [], "v": []}, True),
    # -0.0 and +0.0 are one key value (sort.cc:151 compares with operator<): the second key decides among them
    ("sort_signed_zero", "(sort (order (x DESC) (v ASC)) (all) (scan 0))",
     [[col("x", sp.DOUBLE, [0.0, -0.0, 1.0, -0.0, 0.0, -1.0]), col("v", sp.INT32, [5, 4, 3, 2, 1, 0])]],
     {"x": [1.0, 0.0, -0.0, -0.0, 0.0, -1.0], "v": [3, 1, 2, 4, 5, 0]}, True),
    # cursor/core/merge_union_all_test.cc:131-140 (a1a2b1b2 + a1a3b2b2; the STRING column a/b/c re-coded 1/2/3)
    ("merge_union_all_two_inputs", "(merge_union_all (order (col0 ASC) (col1 ASC)) (scan 0) (scan 1))",
     [[col("col0", sp.INT32, [1, 1, 2, 2]), col("col1", sp.INT32, [1, 2, 1, 2])],
      [col("col0", sp.INT32, [1, 1, 2, 2]), col("col1", sp.INT32, [1, 3, 2, 2])]],
     {"col0": [1, 1, 1, 1, 2, 2, 2, 2], "col1": [1, 1, 2, 3, 1, 2, 2, 2]}, True),
    # :113-129 an empty input on either side
    ("merge_union_all_empty_input", "(merge_union_all (order (col0 ASC) (col1 ASC)) (scan 0) (scan 1) (scan 0))",
     [[col("col0", sp.INT32, []), col("col1", sp.INT32, [])], [col("col0", sp.INT32, [1]), col("col1", sp.INT32, [1])]],
     {"col0": [1], "col1": [1]}, True),
    # :189-219 five one-row inputs in descending / mixed input order
    ("merge_union_all_five_inputs", "(merge_union_all (order (col0 ASC) (col1 ASC)) (scan 0) (scan 1) (scan 2) (scan 3) (scan 4))",
     [[col("col0", sp.INT32, [2]), col("col1", sp.INT32, [2])], [col("col0", sp.INT32, [3]), col("col1", sp.INT32, [3])],
      [col("col0", sp.INT32, [1]), col("col1", sp.INT32, [3])], [col("col0", sp.INT32, [1]), col("col1", sp.INT32, [2])],
      [col("col0", sp.INT32, [1]), col("col1", sp.INT32, [1])]],
     {"col0": [1, 1, 1, 2, 3], "col1": [1, 2, 3, 2, 3]}, True),
    # :262-280 three interleaving inputs; DESC order and a nullable key (NULLs last for DESC): a NOT NULL and a
    # nullable input give a nullable column (merge_union_all.cc:296-311)
    ("merge_union_all_desc_nullable", "(merge_union_all (order (k DESC)) (scan 0) (scan 1) (scan 2))",
     [[col("k", sp.INT64, [9, 5, 1]), col("v", sp.DOUBLE, [0.5, 1.5, 2.5])],
      [ncol("k", sp.INT64, [8, 4, N]), col("v", sp.DOUBLE, [10.0, 11.0, 12.0])],
      [ncol("k", sp.INT64, [7, 6, 2]), col("v", sp.DOUBLE, [20.0, 21.0, 22.0])]],
     {"k": [9, 8, 7, 6, 5, 4, 2, 1, N], "v": [0.5, 10.0, 20.0, 21.0, 1.5, 11.0, 22.0, 2.5, 12.0]}, True),
    ("bound_merge_union_all", "(bound_merge_union_all (order (k ASC)) (scan 0) (scan 1))",
     [[col("k", sp.INT64, [1, 5, 9]), col("v", sp.DOUBLE, [0.5, 1.5, 2.5])],
      [col("k", sp.INT64, [2, 4, 10]), col("v", sp.DOUBLE, [10.0, 11.0, 12.0])]],
     {"k": [1, 2, 4, 5, 9, 10], "v": [0.5, 10.0, 11.0, 1.5, 2.5, 12.0]}, True),
    # cursor/core/aggregate_clusters_test.cc:69-82 shape: three clusters; a key that returns later is a new cluster
    ("aggregate_clusters", "(aggregate_clusters (named k) (aggs (SUM v sum) (COUNT \"\" n)) (scan 0))",
     [[col("k", sp.INT32, [1, 1, 2, 1, 1, 3, 3]), col("v", sp.INT32, [1, 2, 3, 4, 5, 6, 7])]],
     {"k": [1, 2, 1, 3], "sum": [3, 3, 9, 13], "n": [2, 1, 2, 2]}, True),
    # :104-120 no clustering column: one result row
    ("aggregate_clusters_no_key", "(aggregate_clusters (named) (aggs (SUM col0 sum)) (scan 0))",
     [[col("col0", sp.INT32, [13, 3, 7])]], {"sum": [23]}, True),
    # :122-133 empty input with a clustering column
    ("aggregate_clusters_empty", "(aggregate_clusters (named col0) (aggs (SUM col1 sum)) (scan 0))",
     [[col("col0", sp.INT64, []), col("col1", sp.INT32, [])]], {"col0": [], "sum": []}, True),
    # :150-178 three-column key, col1 both clustered by and aggregated (STRING columns re-coded)
    ("aggregate_clusters_three_keys",
     "(aggregate_clusters (rename (col0 A) (col1 B) (col2 C)) (aggs (SUM col1 sum1) (SUM col3 sum3)) (scan 0))",
     [[col("col0", sp.INT64, [1, 1, 1, 1, 1, 1, 1, 1]), col("col1", sp.INT32, [0, 2, 2, 2, 2, 1, 1, 1]),
       col("col2", sp.INT64, [1, 1, 1, 2, 2, 2, 2, 9]), col("col3", sp.INT32, [13, 4, 5, -4, -6, 3, 4, -3])]],
     {"A": [1, 1, 1, 1, 1], "B": [0, 2, 2, 1, 1], "C": [1, 1, 2, 2, 9], "sum1": [0, 4, 4, 2, 1], "sum3": [13, 9, -10, 7, -3]}, True),
    # NULL keys cluster together (NULL equals NULL in the comparator, aggregate_clusters.cc:67-125), all-NULL inputs give NULL
    ("aggregate_clusters_null_keys", "(aggregate_clusters (named k) (aggs (SUM v s) (MIN v m) (COUNT v c)) (scan 0))",
     [[ncol("k", sp.INT64, [N, N, 4, 4, N, 7]), ncol("v", sp.DOUBLE, [1.0, N, N, N, 2.5, 3.0])]],
     {"k": [N, 4, N, 7], "s": [1.0, N, 2.5, 3.0], "m": [1.0, N, 2.5, 3.0], "c": [1, 0, 1, 1]}, True),
    ("bound_aggregate_clusters", "(bound_aggregate_clusters (named k) (aggs (MAX v mx) (LAST v l)) (scan 0))",
     [[col("k", sp.DOUBLE, [0.5, 0.5, -1.0, 0.5]), col("v", sp.INT64, [3, 9, 4, 1])]],
     {"k": [0.5, -1.0, 0.5], "mx": [9, 4, 1], "l": [9, 4, 1]}, True),
    # cursor/core/column_aggregator.cc:333-433 (DistinctAggregator), aggregate_groups_test.cc DISTINCT cases: duplicates of
    # the input inside a group count once, NULL inputs never count; MIN / MAX are unchanged by DISTINCT
    ("group_distinct_count_with_plain_sum", "(group (named k) (aggs (distinct COUNT v c) (SUM v s)) (scan 0))",
     [[col("k", sp.INT32, [1, 3, 1, 3, 1, 1]), ncol("v", sp.INT32, [3, -3, 3, -5, N, 4]), col("w", sp.DOUBLE, [1.5, 2.5, 1.5, 2.5, 1.5, 3.0])]],
     {"k": [1, 3], "c": [2, 2], "s": [10, -8]}, False),
    ("group_distinct_two_inputs", "(group (named k) (aggs (distinct SUM v s) (distinct COUNT w cw)) (scan 0))",
     [[col("k", sp.INT32, [1, 3, 1, 3, 1, 1]), ncol("v", sp.INT32, [3, -3, 3, -5, N, 4]), col("w", sp.DOUBLE, [1.5, 2.5, 1.5, 2.5, 1.5, 3.0])]],
     {"k": [1, 3], "s": [7, -8], "cw": [2, 1]}, False),
    ("scalar_distinct", "(scalar_agg (aggs (distinct COUNT v c) (distinct SUM w sw) (COUNT \"\" n)) (scan 0))",
     [[col("k", sp.INT32, [1, 3, 1, 3, 1, 1]), ncol("v", sp.INT32, [3, -3, 3, -5, N, 4]), col("w", sp.DOUBLE, [1.5, 2.5, 1.5, 2.5, 1.5, 3.0])]],
     {"c": [4], "sw": [7.0], "n": [6]}, True),
    ("group_distinct_min_max", "(group (named k) (aggs (distinct MIN v m) (distinct MAX w x)) (scan 0))",
     [[col("k", sp.INT32, [1, 3, 1, 3, 1, 1]), ncol("v", sp.INT32, [3, -3, 3, -5, N, 4]), col("w", sp.DOUBLE, [1.5, 2.5, 1.5, 2.5, 1.5, 3.0])]],
     {"k": [1, 3], "m": [3, -5], "x": [3.0, 2.5]}, False),
    ("group_distinct_all_null_inputs", "(group (named k) (aggs (distinct COUNT v c) (distinct SUM v s)) (scan 0))",
     [[ncol("k", sp.INT32, [1, N, 1, 2, N]), ncol("v", sp.INT64, [N, 5, N, 7, 5])]],
     {"k": [1, N, 2], "c": [0, 1, 1], "s": [N, 5, 7]}, False),
    # cursor/core/limit.h:27 (limit_test.cc): rows [offset, offset + limit) in the child's order; coalesce.h:30
    # (coalesce_test.cc): the children's columns side by side, the stream ends with its shortest input
    ("limit_middle", "(limit 2 5 (scan 0))",
     [[col("a", sp.INT32, list(range(10))), ncol("s", sp.STRING, ["x", N, "yy", "z", "", "q", N, "w", "e", "r"])]],
     {"a": [2, 3, 4, 5, 6], "s": [b"yy", b"z", b"", b"q", N]}, True),
    ("limit_past_the_end", "(limit 9 5 (scan 0))",
     [[col("a", sp.INT32, list(range(10))), ncol("s", sp.STRING, ["x", N, "yy", "z", "", "q", N, "w", "e", "r"])]],
     {"a": [9], "s": [b"r"]}, True),
    ("limit_nothing", "(limit 20 5 (scan 0))",
     [[col("a", sp.INT32, list(range(10))), ncol("s", sp.STRING, ["x", N, "yy", "z", "", "q", N, "w", "e", "r"])]],
     {"a": [], "s": []}, True),
    ("limit_over_sort", "(limit 1 3 (sort (order (a DESC)) (all) (scan 0)))",
     [[col("a", sp.INT32, [5, 1, 9, 7, 3])]],
     {"a": [7, 5, 3]}, True),
    ("coalesce_two", "(limit 1 3 (coalesce (scan 0) (compute (as c (plus (col b) (i64 1))) (scan 1))))",
     [[col("a", sp.INT32, list(range(10)))], [col("b", sp.INT64, [i * 10 for i in range(10)])]],
     {"a": [1, 2, 3], "c": [11, 21, 31]}, True),
    ("coalesce_shortest_input", "(coalesce (scan 0) (scan 1))",
     [[col("b", sp.INT64, [i * 10 for i in range(10)])], [col("a", sp.INT64, [1, 2, 3])]],
     {"b": [0, 10, 20], "a": [1, 2, 3]}, True),
]

# STRING / BINARY columns: known-answer vectors of the reference's tests with their own STRING cells.
S = sp.STRING
GOLDEN_STRINGS = [
    # cursor/core/aggregate_groups_test.cc:379-401 (GroupByTwoColumns)
    ("str_group_two_columns", "(group (named col0 col1) (aggs (SUM col2 sum)) (scan 0))",
     [[col("col0", S, ["foo", "bar", "foo", "bar"]), col("col1", sp.INT32, [1, 2, 1, 3]), col("col2", sp.INT32, [3, -3, 4, -5])]],
     {"col0": ["foo", "bar", "bar"], "col1": [1, 2, 3], "sum": [7, -3, -5]}, False),
    # cursor/core/sort_test.cc:189-211 (OneStringColumnWithDuplicatesAndNulls)
    ("str_sort_duplicates_and_nulls", "(sort (order (col0 ASC)) (all) (scan 0))",
     [[ncol("col0", S, ["a", "c", "a", N, "d", N, "e"])]],
     {"col0": [N, N, "a", "a", "c", "d", "e"]}, True),
    # :213-237 (OneIntegerColumnMostlyNullsDescending): a STRING payload follows the permutation
    ("str_sort_payload", "(sort (order (col0 DESC)) (all) (scan 0))",
     [[ncol("col0", sp.INT32, [N, N, N, 7, N, N, N]), col("col1", S, ["a", "a", "a", "g", "a", "a", "a"])]],
     {"col0": [7, N, N, N, N, N, N], "col1": ["g", "a", "a", "a", "a", "a", "a"]}, True),
    # cursor/core/hash_join_test.cc:60-160: (INT64, STRING) rows joined on the STRING column, both key kinds
    ("str_join_inner_unique",
     "(hash_join INNER (named col1) (named col1) (multi (0 (rename (col0 L.col0) (col1 L.col1))) (1 (rename (col0 R.col0) (col1 R.col1)))) UNIQUE (scan 0) (scan 1))",
     [[col("col0", sp.INT64, [1, 2, 3, 4, 5]), col("col1", S, ["a", "b", "c", "d", "e"])],
      [col("col0", sp.INT64, [6, 5, 4, 3, 2, 1]), col("col1", S, ["f", "e", "d", "c", "b", "a"])]],
     {"L.col0": [1, 2, 3, 4, 5], "L.col1": ["a", "b", "c", "d", "e"], "R.col0": [1, 2, 3, 4, 5], "R.col1": ["a", "b", "c", "d", "e"]}, True),
    ("str_join_not_unique_2b2b2c",
     "(hash_join INNER (named col0) (named col0) (multi (0 (rename (col0 L.col0) (col1 L.col1))) (1 (rename (col0 R.col0) (col1 R.col1)))) NOT_UNIQUE (scan 0) (scan 0))",
     [[col("col0", sp.INT64, [2, 2, 2]), col("col1", S, ["b", "b", "c"])]],
     {"L.col0": [2] * 9, "L.col1": ["b", "b", "b", "b", "b", "b", "c", "c", "c"], "R.col0": [2] * 9,
      "R.col1": ["b", "b", "c", "b", "b", "c", "b", "b", "c"]}, True),
    # NULL STRING keys never match (builder_1a1NNaNN_), LEFT_OUTER keeps the lhs rows
    ("str_join_left_outer_null_keys",
     "(hash_join LEFT_OUTER (named col1) (named col1) (multi (0 (rename (col0 L.col0) (col1 L.col1))) (1 (rename (col0 R.col0)))) NOT_UNIQUE (scan 0) (scan 0))",
     [[ncol("col0", sp.INT64, [1, 1, N, N]), ncol("col1", S, ["a", N, "a", N])]],
     {"L.col0": [1, 1, 1, N, N, N], "L.col1": ["a", "a", N, "a", "a", N], "R.col0": [1, N, N, 1, N, N]}, True),
    # cursor/core/aggregate_clusters_test.cc:150-178 with its own STRING columns
    ("str_aggregate_clusters_three_keys",
     "(aggregate_clusters (rename (col0 A) (col1 B) (col2 C)) (aggs (SUM col1 sum1) (SUM col3 sum3)) (scan 0))",
     [[col("col0", S, ["a"] * 8), col("col1", sp.INT32, [0, 2, 2, 2, 2, 1, 1, 1]),
       col("col2", S, ["a", "a", "a", "b", "b", "b", "b", "bbbbbbbb"]), col("col3", sp.INT32, [13, 4, 5, -4, -6, 3, 4, -3])]],
     {"A": ["a"] * 5, "B": [0, 2, 2, 1, 1], "C": ["a", "a", "b", "b", "bbbbbbbb"], "sum1": [0, 4, 4, 2, 1], "sum3": [13, 9, -10, 7, -3]}, True),
    # cursor/core/merge_union_all_test.cc:131-140 with its own STRING column
    ("str_merge_union_all", "(merge_union_all (order (col0 ASC) (col1 ASC)) (scan 0) (scan 1))",
     [[col("col0", S, ["a", "a", "b", "b"]), col("col1", sp.INT32, [1, 2, 1, 2])],
      [col("col0", S, ["a", "a", "b", "b"]), col("col1", sp.INT32, [1, 3, 2, 2])]],
     {"col0": ["a", "a", "a", "a", "b", "b", "b", "b"], "col1": [1, 1, 2, 3, 1, 2, 2, 2]}, True),
    # expression/core/comparison_expressions_test.cc: STRING comparisons (memcmp order, the shorter one first), NULL operands
    ("str_compare_literal", "(compute (compound (as eq (equal (col s) (str \"bob\"))) (as lt (less (col s) (str \"bob\"))) "
                            "(as ge (greater_or_equal (col s) (str \"bo\"))) (as ne (not_equal (col s) (str \"\")))) (scan 0))",
     [[ncol("s", S, ["bob", "alice", "", N, "bo", "bobby", "Bob"])]],
     {"eq": [True, False, False, N, False, False, False], "lt": [False, True, True, N, True, False, True],
      "ge": [True, False, False, N, True, True, False], "ne": [True, True, False, N, True, True, True]}, True),
    ("str_compare_columns", "(filter (less_or_equal (col a) (col b)) (all) (scan 0))",
     [[col("a", S, ["x", "abc", "ab", "b", ""]), ncol("b", S, ["x", "ab", "abc", N, "a"]), col("v", sp.INT32, [1, 2, 3, 4, 5])]],
     {"a": ["x", "ab", ""], "b": ["x", "abc", "a"], "v": [1, 3, 5]}, True),
    # MIN / MAX / FIRST / LAST / COUNT over a STRING column (column_aggregator.cc:108-166,314-377); all-NULL group -> NULL
    ("str_aggregates", "(group (named k) (aggs (MIN s mn) (MAX s mx) (FIRST s f) (LAST s l) (COUNT s c)) (scan 0))",
     [[col("k", sp.INT32, [1, 2, 1, 2, 1, 3]), ncol("s", S, ["pear", "fig", "apple", N, "plum", N])]],
     {"k": [1, 2, 3], "mn": ["apple", "fig", N], "mx": ["plum", "fig", N], "f": ["pear", "fig", N], "l": ["plum", "fig", N], "c": [3, 1, 0]}, False),
    ("str_scalar_aggregate", "(scalar_agg (aggs (MIN s mn) (MAX s mx) (COUNT \"\" n)) (scan 0))",
     [[col("s", S, ["pear", "fig", "apple", "plum"])]], {"mn": ["apple"], "mx": ["plum"], "n": [4]}, True),
    # BINARY cells may hold zero bytes: "ab" < "ab\0" < "ab\0\0" < "ab\1"; cells longer than one 8-byte ranking round
    ("bin_sort_embedded_zero", "(sort (order (b ASC)) (all) (scan 0))",
     [[col("b", sp.BINARY, [b"ab\x01", b"ab\x00\x00", b"ab", b"ab\x00", b"0123456789abcdefX", b"0123456789abcdef", b"0123456789abcdeg"])]],
     {"b": [b"0123456789abcdef", b"0123456789abcdefX", b"0123456789abcdeg", b"ab", b"ab\x00", b"ab\x00\x00", b"ab\x01"]}, True),
    ("str_is_null_and_project", "(filter (not (is_null (col s))) (named s) (compute (compound (col s) (as n (plus (col v) (i32 1)))) (scan 0)))",
     [[ncol("s", S, ["q", N, "w"]), col("v", sp.INT32, [1, 2, 3])]], {"s": ["q", "w"]}, True),
    ("str_empty_table", "(group (named s) (aggs (COUNT \"\" n)) (scan 0))", [[col("s", S, [])]], {"s": [], "n": []}, True),
    # expression/core/elementary_expressions.h:36-46 ParseStringQuiet / ParseStringNulling over literals (the form
    # test/guide/join.cc:316-329 uses); parsers of base/infrastructure/types_infrastructure.cc:154-258: whitespace at
    # either end accepted, " %Y/%m/%d " dates in days since 1970 (earlier dates refused), invalid input -> NULL
    ("parse_string_constants",
     "(compute (compound (as d (parse_string_nulling DATE (str \"1991/01/01\"))) (as bad (parse_string_nulling DATE (str \"Mort\")))"
     " (as old (parse_string_nulling DATE (str \"1929/01/01\"))) (as i (parse_string_quiet INT32 (str \" 42 \")))"
     " (as u (parse_string_nulling UINT32 (str \"-1\"))) (as f (parse_string_nulling DOUBLE (str \"2.5\")))"
     " (as b (parse_string_nulling BOOL (str \"Yes\"))) (as t (parse_string_nulling DATETIME (str \"2001/02/03-04:05:06\")))"
     " (as k (col k))) (scan 0))",
     [[col("k", sp.INT32, [1, 2])]],
     {"d": [7670, 7670], "bad": [N, N], "old": [N, N], "i": [42, 42], "u": [N, N], "f": [2.5, 2.5], "b": [True, True],
      "t": [981173106000000, 981173106000000], "k": [1, 2]}, True),
]


def check_result(r, expected, ordered):
    """Compares a PlanResult with {name: list}; NULL cells compare by is_null only
    (testing/view_comparator.cc:54-61); unordered results are sorted on all columns first."""
    assert r.code == 0, (r.code, r.error)
    assert r.names == list(expected.keys()), (r.names, list(expected.keys()))
    n = len(next(iter(expected.values()))) if expected else 0
    assert r.rows == n, (r.rows, n)
    got_rows, want_rows = [], []
    for i in range(n):
        g, w = [], []
        for j, name in enumerate(r.names):
            isn = bool(r.nulls[j][i]) if r.nulls[j] is not None else False
            v = r.columns[j][i]
            v = v.item() if hasattr(v, "item") else v
            g.append(("null",) if isn else ("v", _norm(v)))
            e = expected[name][i]
            w.append(("null",) if e is None else ("v", _norm(e)))
        got_rows.append(tuple(g))
        want_rows.append(tuple(w))
    if not ordered:
        got_rows.sort(key=repr)
        want_rows.sort(key=repr)
    assert got_rows == want_rows, (got_rows, want_rows)


def _norm(v):
    if isinstance(v, float):
        if v != v:
            return "nan"
        return float(v)
    if isinstance(v, bool):
        return bool(v)
    if isinstance(v, str):
        return v.encode()      # STRING / BINARY cells compare as bytes
    return v


def same_results(a, b, ordered=True, sort_cols=None):
    """Bit-exact comparison of two PlanResults (masking data under NULL)."""
    assert a.code == b.code, (a.code, a.error, b.code, b.error)
    if a.code != 0:
        return
    assert a.names == b.names and a.dtypes == b.dtypes and a.nullable == b.nullable, \
        (a.names, b.names, a.dtypes, b.dtypes, a.nullable, b.nullable)
    assert a.rows == b.rows, (a.rows, b.rows)
    if a.rows == 0:
        return
    ca, cb = _masked(a), _masked(b)
    if not ordered:
        ca, cb = _sorted(ca, sort_cols), _sorted(cb, sort_cols)
    for j in range(len(ca)):
        va, na = ca[j]
        vb, nb = cb[j]
        assert np.array_equal(na, nb), "null vectors differ in column %s" % a.names[j]
        if va.dtype == object:
            assert list(va) == list(vb), "column %s differs: %s vs %s" % (a.names[j], va[:8], vb[:8])
            continue
        assert np.array_equal(va.view(np.uint8), vb.view(np.uint8)), \
            "column %s differs: %s vs %s" % (a.names[j], va[:8], vb[:8])


def _masked(r):
    out = []
    for j in range(len(r.columns)):
        v = r.columns[j].copy()
        n = r.nulls[j] if r.nulls[j] is not None else np.zeros(r.rows, dtype=np.bool_)
        if v.dtype == np.bool_:
            v = v.astype(np.uint8)
        if v.dtype == object:
            v[n] = b""
            out.append((v, n))
            continue
        v[n] = 0
        if v.dtype.kind == "f":
            v = v + 0.0   # canonical zero sign is not touched; NaN payloads are kept bit-exact
        out.append((v, n))
    return out


def _sorted(cols, sort_cols):
    keys = []
    idx = list(range(len(cols))) if sort_cols is None else sort_cols
    for j in reversed(idx):
        v, n = cols[j]
        if v.dtype == object:   # STRING / BINARY: dense ranks of the cells
            keys.append(np.unique(np.array([bytes(x) for x in v], dtype=object), return_inverse=True)[1].astype(np.int64))
            keys.append(n)
            continue
        keys.append(v.view(np.uint64) if v.dtype.itemsize == 8 else v.astype(np.int64))
        keys.append(n)
    order = np.lexsort(keys)
    return [(v[order], n[order]) for v, n in cols]


# ---------------------------------------------------------------------------------------------
# The Bound* factories (cursor/core/compute.h:36, filter.h:43, project.h:34, scan_view.h:52,
# aggregate.h:254,345, sort.h:114): each pair is the same plan through the Operation factories and
# through the bound ones; both must give the same result, on the oracle and on the GPU.
def bound_tables():
    import numpy as np
    from supersonic_b200 import ssplan as sp
    rng = np.random.default_rng(99)
    n = 20_000
    return [[sp.Column("a", sp.INT64, rng.integers(-1000, 1000, n), is_null=rng.random(n) < 0.05),
             sp.Column("b", sp.INT64, rng.integers(0, 50, n)),
             sp.Column("x", sp.DOUBLE, rng.integers(0, 4096, n) / 8.0),
             sp.Column("k", sp.INT32, rng.integers(0, 7, n).astype(np.int32))]]


BOUND_PAIRS = [
    ("compute",
     "(compute (compound (as s (plus (col a) (col b))) (col x)) (scan 0))",
     "(bound_compute (compound (as s (plus (col a) (col b))) (col x)) (bound_scan 0))", True),
    ("filter",
     "(filter (less (col b) (i64 10)) (named a x) (scan 0))",
     "(bound_filter (less (col b) (i64 10)) (named a x) (bound_scan 0))", True),
    ("project",
     "(project (named x k) (scan 0))",
     "(bound_project (named x k) (bound_scan 0))", True),
    ("filter_over_compute",
     "(filter (greater (col s) (i64 0)) (all) (compute (compound (as s (plus (col a) (col b))) (col k)) (scan 0)))",
     "(bound_filter (greater (col s) (i64 0)) (all) (bound_compute (compound (as s (plus (col a) (col b))) (col k)) (scan 0)))", True),
    ("group",
     "(group (named k) (aggs (SUM x sx) (COUNT a ca) (MIN a mn) (COUNT \"\" n)) (scan 0))",
     "(bound_group (named k) (aggs (SUM x sx) (COUNT a ca) (MIN a mn) (COUNT \"\" n)) (bound_scan 0))", False),
    ("scalar_agg",
     "(scalar_agg (aggs (SUM x sx) (MAX a mx) (COUNT \"\" n)) (scan 0))",
     "(bound_scalar_agg (aggs (SUM x sx) (MAX a mx) (COUNT \"\" n)) (bound_scan 0))", True),
    ("sort",
     "(sort (order (k ASC) (x DESC) (b ASC) (a ASC)) (all) (scan 0))",
     "(bound_sort (order (k ASC) (x DESC) (b ASC) (a ASC)) (all) (bound_scan 0))", False),
    ("group_over_bound_filter",
     "(group (named k) (aggs (SUM b sb)) (filter (less (col b) (i64 25)) (all) (scan 0)))",
     "(bound_group (named k) (aggs (SUM b sb)) (bound_filter (less (col b) (i64 25)) (all) (bound_scan 0)))", False),
]


# ------------------------------------------------------------------------------------------------
# Signaling operators under skip vectors: the reference evaluates a sub-expression only on
# the rows its parent leaves (the taken branch of IF / CASE, the undecided side of AND / OR, IFNULL's
# substitute where the value is NULL, rows whose earlier operand is not NULL: elementary_bound_expressions.cc:
# 262-327,406-539,896-1050, abstract_bound_expressions.h:129-147) and a Compute above a Filter only on the
# kept rows (filter.cc:96-128); a signaling division fails only there (binary_column_computers.h:137-166).
# Expected return codes were taken from the oracle (tests/test_oracle_golden.py pins them on the CPU).
def signaling_tables():
    import numpy as np
    from supersonic_b200 import ssplan as sp
    n = 5000
    rng = np.random.default_rng(1)
    a = rng.integers(-100, 100, n).astype(np.int32)
    b = rng.integers(-3, 4, n).astype(np.int32)          # zeros inside
    nx_null = (b == 0) | (rng.random(n) < 0.1)
    cols = [sp.Column("a", sp.INT32, a), sp.Column("b", sp.INT32, b),
            sp.Column("nx", sp.INT32, rng.integers(0, 9, n).astype(np.int32), is_null=nx_null),      # NULL wherever b == 0
            sp.Column("ny", sp.INT32, rng.integers(0, 9, n).astype(np.int32), is_null=(b != 0)),      # NULL wherever b != 0
            sp.Column("k", sp.INT64, rng.integers(0, 5, n))]
    return [cols]


_DIV = "(cpp_divide_signaling (col a) (col b))"
SIGNALING_CASES = [
    ("filter_below_compute", "(compute (as q %s) (filter (not_equal (col b) (i32 0)) (all) (scan 0)))" % _DIV, 0),
    ("filter_below_compute_fails", "(compute (as q %s) (filter (not_equal (col a) (i32 1000)) (all) (scan 0)))" % _DIV, 104),
    ("compute_below_filter_fails", "(filter (not_equal (col b) (i32 0)) (all) (compute (compound (col b) (as q %s)) (scan 0)))" % _DIV, 104),
    ("if_untaken_branch", "(compute (as q (if (equal (col b) (i32 0)) (i32 0) %s)) (scan 0))" % _DIV, 0),
    ("if_taken_branch_fails", "(compute (as q (if (not_equal (col b) (i32 0)) (i32 0) %s)) (scan 0))" % _DIV, 104),
    ("if_then_side", "(compute (as q (if (not_equal (col b) (i32 0)) %s (i32 7))) (scan 0))" % _DIV, 0),
    ("nulling_if_null_condition", "(compute (as q (nulling_if (less (col nx) (i32 100)) %s (i32 7))) (scan 0))" % _DIV, 0),
    ("nulling_if_else_fails", "(compute (as q (nulling_if (less (col ny) (i32 -1)) (i32 7) %s)) (scan 0))" % _DIV, 104),
    ("case_then", "(compute (as q (case (col b) (i32 -1) (i32 0) (i32 0) (i32 1) %s (i32 2) %s)) (scan 0))" % (_DIV, _DIV), 0),
    ("case_else", "(compute (as q (case (col b) %s (i32 0) (i32 0))) (scan 0))" % _DIV, 0),
    ("case_else_fails", "(compute (as q (case (col b) %s (i32 5) (i32 0))) (scan 0))" % _DIV, 104),
    ("and_short_circuit", "(compute (as q (and (not_equal (col b) (i32 0)) (greater %s (i32 1)))) (scan 0))" % _DIV, 0),
    ("and_left_side_fails", "(compute (as q (and (greater %s (i32 1)) (not_equal (col b) (i32 0)))) (scan 0))" % _DIV, 104),
    ("or_short_circuit", "(compute (as q (or (equal (col b) (i32 0)) (greater (modulus_signaling (col a) (col b)) (i32 0)))) (scan 0))", 0),
    ("and_not_short_circuit", "(compute (as q (and_not (equal (col b) (i32 0)) (greater %s (i32 1)))) (scan 0))" % _DIV, 0),
    ("if_null_substitute", "(compute (as q (if_null (col ny) %s)) (scan 0))" % _DIV, 0),
    ("if_null_substitute_fails", "(compute (as q (if_null (col nx) %s)) (scan 0))" % _DIV, 104),
    ("null_left_operand_skips_right", "(compute (as q (plus (col nx) %s)) (scan 0))" % _DIV, 0),
    ("null_right_operand_does_not_skip_left", "(compute (as q (plus %s (col nx))) (scan 0))" % _DIV, 104),
    ("nested", "(compute (as q (if (equal (col b) (i32 0)) (i32 -1) (plus (i32 1) (if (greater (col a) (i32 0)) %s (negate %s))))) (scan 0))" % (_DIV, _DIV), 0),
    ("filter_over_filter", "(filter (greater %s (i32 0)) (all) (filter (not_equal (col b) (i32 0)) (all) (scan 0)))" % _DIV, 0),
    ("group_over_compute_over_filter", "(group (named k) (aggs (SUM q s) (COUNT \"\" n)) (compute (compound (col k) (as q (cast INT64 %s))) (filter (not_equal (col b) (i32 0)) (all) (scan 0))))" % _DIV, 0),
    ("divide_signaling_double", "(compute (as q (if (equal (col b) (i32 0)) (f64 0) (divide_signaling (col a) (col b)))) (scan 0))", 0),
    ("in_list", "(compute (as q (in (col a) (i32 3) (if (equal (col b) (i32 0)) (i32 0) %s))) (scan 0))" % _DIV, 0),
]


# ------------------------------------------------------------------------------------------------
# Expression::Bind + BoundExpressionTree::Evaluate (expression.cc:41-94), the entry point of
# test/guide/primer.cc. Same tuple layout as GOLDEN; `code` != 0 expects that failure.
EVALUATE_CASES = [
    # test/guide/primer.cc:103-136,209-222 (PrimerExample1.ColumnAddTest): Plus(AttributeAt(0), AttributeAt(1)) bound
    # with a capacity of 2048 rows over a = 0..7, b = {3,4,6,8,1,2,2,9}
    ("primer_column_add", "(evaluate (plus (at 0) (at 1)) 0 2048)",
     [[col("a", sp.INT32, [0, 1, 2, 3, 4, 5, 6, 7]), col("b", sp.INT32, [3, 4, 6, 8, 1, 2, 2, 9])]],
     {"(a + b)": [3, 5, 8, 11, 5, 7, 8, 16]}, 0),
    # several calls on one bound tree (three slices of at most three rows), two result columns, a NULL cell
    ("evaluate_in_slices", "(evaluate (compound (as s (plus (col a) (col b))) (as m (multiply (col a) (i32 2)))) 0 3)",
     [[col("a", sp.INT32, [1, 2, 3, 1, 2, 3, 1, 2]), ncol("b", sp.DOUBLE, [1.5, 3.0, N, 7.6, 5.5, 2.0, 1.6, 9.5])]],
     {"s": [2.5, 5.0, N, 8.6, 7.5, 5.0, 2.6, 11.5], "m": [2, 4, 6, 2, 4, 6, 2, 4]}, 0),
    # expression.cc:57-66: a view larger than the capacity the tree was bound for
    ("evaluate_too_many_rows", "(evaluate (plus (col a) (col b)) 0 4 5)",
     [[col("a", sp.INT32, [0, 1, 2, 3, 4, 5, 6, 7]), col("b", sp.INT32, [3, 4, 6, 8, 1, 2, 2, 9])]], {}, 302),
    ("evaluate_signaling_failure", "(evaluate (cpp_divide_signaling (col a) (minus (col a) (i32 1))) 0 1024)",
     [[col("a", sp.INT32, [0, 1, 2, 3])]], {}, 104),
    ("evaluate_empty_view", "(evaluate (plus (col a) (col a)) 0 16)", [[col("a", sp.INT64, [])]], {"(a + a)": []}, 0),
]
