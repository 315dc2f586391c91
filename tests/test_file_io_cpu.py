"""The reference's block format on disk (cursor/infrastructure/file_io.cc:70-420): files written by the
mirror's FileOutput are byte-identical to the reference's, and each side scans the other's files with FileInput.
Host code on both sides, so the whole check runs without a GPU (the source is a ViewCursor over the test table)."""
import os

import numpy as np
import pytest

from cases import N, col, ncol, same_results
from supersonic_b200 import ssplan as sp


def _tables(rows):
    rng = np.random.default_rng(rows)
    words = ["", "a", "Supersonic", "B200", "x" * 40, "columnar", "query"]
    strs = [words[i] for i in rng.integers(0, len(words), rows)]
    nulls = list(rng.random(rows) < 0.2)
    return [[col("i", sp.INT32, list(map(int, rng.integers(-1000, 1000, rows)))),
             ncol("d", sp.DOUBLE, [None if n else float(v) for n, v in zip(nulls, rng.integers(0, 1 << 20, rows) / 8.0)]),
             ncol("s", sp.STRING, [None if n else w for n, w in zip(nulls[::-1], strs)]),
             col("b", sp.BOOL, [bool(v) for v in rng.integers(0, 2, rows)]),
             col("t", sp.STRING, strs),
             ncol("u", sp.UINT64, [None if n else int(v) for n, v in zip(nulls, rng.integers(0, 1 << 62, rows))])]]


@pytest.mark.parametrize("rows", [1, 5, 8192, 8193, 20000])
def test_files_are_byte_identical_and_cross_readable(ref, b200, rows, tmp_path):
    tables = _tables(rows)
    fr, fb = str(tmp_path / "ref.ssb"), str(tmp_path / "b200.ssb")
    a = ref.run("(file_write %s (scan 0))" % fr, tables)
    b = b200.run("(file_write %s (scan 0))" % fb, tables)
    assert a.code == 0 and b.code == 0, (a.code, a.error, b.code, b.error)
    same_results(a, b)
    assert a.rows == rows
    assert open(fr, "rb").read() == open(fb, "rb").read()
    # each implementation scans the other's file (chunks of at most 8192 rows, Next(1) .. Next(default))
    for next_rows in (0, 1, 1000):
        same_results(a, b200.run("(file_read %s 0)" % fr, tables, next_max_rows=next_rows))
        same_results(a, ref.run("(file_read %s 0)" % fb, tables, next_max_rows=next_rows))


def test_truncated_file_fails_like_the_reference(ref, b200, tmp_path):
    tables = _tables(100)
    path = str(tmp_path / "t.ssb")
    assert b200.run("(file_write %s (scan 0))" % path, tables).code == 0
    data = open(path, "rb").read()
    open(path, "wb").write(data[:len(data) // 2])
    a, b = ref.run("(file_read %s 0)" % path, tables), b200.run("(file_read %s 0)" % path, tables)
    assert a.code == b.code == 101, (a.code, b.code)   # ERROR_GENERAL_IO_ERROR
    open(path, "wb").write(b"")
    a, b = ref.run("(file_read %s 0)" % path, tables), b200.run("(file_read %s 0)" % path, tables)
    assert a.code == b.code == 0 and a.rows == b.rows == 0
