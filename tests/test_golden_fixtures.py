"""Committed golden fixtures (tests/golden/*.npz, generated from the oracle by tests/golden/make_golden.py):
the oracle must still reproduce them (CPU), and the B200 path must produce them (GPU)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402


def _canonical(names, dtypes, rows, cols, nulls, ordered):
    cols = [np.where(n, 0, c.astype(np.uint8) if c.dtype == np.bool_ else c) for c, n in zip(cols, nulls)]
    if not ordered and rows > 1:
        order = np.lexsort([x for c, n in reversed(list(zip(cols, nulls))) for x in (c, n)])
        cols = [c[order] for c in cols]
        nulls = [n[order] for n in nulls]
    return cols, nulls


def _check(result, name, ordered):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    assert result.code == 0, result.error
    assert list(g["names"]) == result.names and list(g["dtypes"]) == result.dtypes
    rows = int(g["rows"])
    assert result.rows == rows
    want_nulls = [np.unpackbits(g["null%d" % j])[:rows].astype(bool) for j in range(len(result.names))]
    want_cols = [g["col%d" % j] for j in range(len(result.names))]
    got_nulls = [n if n is not None else np.zeros(rows, dtype=bool) for n in result.nulls]
    wc, wn = _canonical(result.names, result.dtypes, rows, want_cols, want_nulls, ordered)
    gc, gn = _canonical(result.names, result.dtypes, rows, result.columns, got_nulls, ordered)
    for j in range(len(result.names)):
        assert np.array_equal(wn[j], gn[j]), (name, result.names[j], "NULL pattern")
        assert np.array_equal(wc[j], gc[j]), (name, result.names[j], "values")


@pytest.mark.parametrize("name", sorted(make_golden.PLANS))
def test_oracle_reproduces_golden_fixtures(ref, name):
    plan, ordered = make_golden.PLANS[name]
    _check(ref.run(plan, make_golden.tables()), name, ordered)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(make_golden.PLANS))
@pytest.mark.parametrize("next_rows", [0, 777])
def test_b200_reproduces_golden_fixtures(b200, name, next_rows):
    plan, ordered = make_golden.PLANS[name]
    _check(b200.run(plan, make_golden.tables(), next_max_rows=next_rows), name, ordered)
