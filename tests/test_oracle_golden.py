"""Pins the oracle (the unmodified reference built against shims, oracle/_ref) to the
known-answer vectors of the reference's own tests. CPU only."""
import pytest

from cases import GOLDEN, check_result


@pytest.mark.parametrize("case", GOLDEN, ids=[c[0] for c in GOLDEN])
@pytest.mark.parametrize("next_rows", [0, 1, 2])
def test_oracle_reproduces_reference_vectors(ref, case, next_rows):
    _, plan, tables, expected, ordered = case
    r = ref.run(plan, tables, next_max_rows=next_rows)
    check_result(r, expected, ordered)
