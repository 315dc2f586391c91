"""Pins the oracle (the unmodified reference built against shims, oracle/_ref) to the
known-answer vectors of the reference's own tests. CPU only."""
import pytest

from cases import GOLDEN, GOLDEN_LATE, GOLDEN_STRINGS, check_result

ALL_GOLDEN = GOLDEN + GOLDEN_LATE + GOLDEN_STRINGS


@pytest.mark.parametrize("case", ALL_GOLDEN, ids=[c[0] for c in ALL_GOLDEN])
@pytest.mark.parametrize("next_rows", [0, 1, 2])
def test_oracle_reproduces_reference_vectors(ref, case, next_rows):
    _, plan, tables, expected, ordered = case
    r = ref.run(plan, tables, next_max_rows=next_rows)
    check_result(r, expected, ordered)


from cases import BOUND_PAIRS, bound_tables, same_results  # noqa: E402


@pytest.mark.parametrize("pair", BOUND_PAIRS, ids=[p[0] for p in BOUND_PAIRS])
def test_oracle_bound_factories_equal_operation_factories(ref, pair):
    """The plan driver's bound_* verbs (BoundCompute, BoundFilter, ... of the reference) give what
    the Operation factories give: pins the driver code the GPU parity tests then reuse."""
    name, unbound, bound, ordered = pair
    tables = bound_tables()
    a, b = ref.run(unbound, tables), ref.run(bound, tables)
    assert a.code == 0 and b.code == 0, (a.error, b.error)
    same_results(a, b, ordered=ordered, sort_cols=None if ordered else list(range(len(a.columns))))


from cases import SIGNALING_CASES, signaling_tables  # noqa: E402


@pytest.mark.parametrize("case", SIGNALING_CASES, ids=[c[0] for c in SIGNALING_CASES])
def test_oracle_signaling_ops_follow_skip_vectors(ref, case):
    """The return codes the GPU parity test expects are the reference's own: a signaling division fails
    only on rows its skip vector leaves (see cases.py)."""
    _, plan, code = case
    r = ref.run(plan, signaling_tables())
    assert r.code == code, (r.code, r.error)


from cases import EVALUATE_CASES  # noqa: E402


@pytest.mark.parametrize("case", EVALUATE_CASES, ids=[c[0] for c in EVALUATE_CASES])
def test_oracle_bound_expression_tree_evaluate(ref, case):
    """BoundExpressionTree::Evaluate through the plan driver's (evaluate ...) verb: primer.cc's golden values."""
    _, plan, tables, expected, code = case
    r = ref.run(plan, tables)
    assert r.code == code, (r.code, r.error)
    if code == 0:
        check_result(r, expected, True)
