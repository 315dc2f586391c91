"""ctypes binding of the plan driver (supersonic_b200/plan_driver/ssplan.h).

The same C interface is exported by two libraries:
  * supersonic_b200/lib/libssb200_plan.so -- the B200 product (host C++ mirror of
    supersonic.h on top of the CUDA C-ABI library)
  * oracle/_ref/libssref.so               -- the unmodified reference (test oracle)
This module is only plumbing: numpy arrays in, numpy arrays out.
"""
import ctypes as C
import os

import numpy as np

# supersonic::DataType numbers (supersonic/proto/supersonic.proto:15-37)
INT32, INT64, UINT64, DATETIME, DOUBLE, BOOL, UINT32, FLOAT, DATE, ENUM = 1, 2, 3, 4, 5, 6, 8, 9, 10, 13
STRING, BINARY = 0, 7
VARLEN = (STRING, BINARY)

NP_OF = {
    INT32: np.int32, INT64: np.int64, UINT32: np.uint32, UINT64: np.uint64,
    FLOAT: np.float32, DOUBLE: np.float64, BOOL: np.bool_, DATE: np.int32,
    DATETIME: np.int64, ENUM: np.int32,
}
NAME_OF = {INT32: "INT32", INT64: "INT64", UINT32: "UINT32", UINT64: "UINT64", FLOAT: "FLOAT",
           DOUBLE: "DOUBLE", BOOL: "BOOL", DATE: "DATE", DATETIME: "DATETIME", ENUM: "ENUM", STRING: "STRING",
           BINARY: "BINARY"}
DTYPE_OF_NAME = {v: k for k, v in NAME_OF.items()}

SSPLAN_DISCARD = 1
SSPLAN_BIND_ONLY = 2
SSPLAN_SPY = 4

OK = 0
ERROR_MEMORY_EXCEEDED = 102
ERROR_NOT_IMPLEMENTED = 103
ERROR_EVALUATION_ERROR = 104


class _Column(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_int32), ("nullable", C.c_int32),
                ("data", C.c_void_p), ("is_null", C.c_void_p)]


class _Table(C.Structure):
    _fields_ = [("ncols", C.c_int32), ("rows", C.c_int64), ("cols", C.POINTER(_Column))]


class Column(object):
    """One input column: name, supersonic dtype number, values, optional is_null (bool array)."""

    def __init__(self, name, dtype, data, is_null=None, nullable=None):
        self.name = name
        self.dtype = dtype
        if dtype in VARLEN:
            # STRING / BINARY: python str / bytes values -> StringPiece cells { const char* ptr; int64 length }
            # pointing into one bytes buffer (kept alive by this object)
            self.values = [v.encode() if isinstance(v, str) else bytes(v) for v in data]
            self._buf = np.frombuffer(b"".join(self.values) + b"\0", dtype=np.uint8).copy()
            lens = np.array([len(v) for v in self.values], dtype=np.int64)
            cells = np.zeros((len(self.values), 2), dtype=np.int64)
            if len(self.values):
                cells[:, 0] = self._buf.ctypes.data + np.concatenate(([0], np.cumsum(lens)[:-1]))
                cells[:, 1] = lens
            self.data = cells
        else:
            self.data = np.ascontiguousarray(data, dtype=NP_OF[dtype])
        self.is_null = None if is_null is None else np.ascontiguousarray(is_null, dtype=np.bool_)
        self.nullable = (is_null is not None) if nullable is None else nullable


class PlanResult(object):
    def __init__(self, code, error, names, dtypes, nullable, columns, nulls, rows,
                 create_s, drain_s, next_calls):
        self.code, self.error = code, error
        self.names, self.dtypes, self.nullable = names, dtypes, nullable
        self.columns, self.nulls, self.rows = columns, nulls, rows
        self.create_seconds, self.drain_seconds, self.next_calls = create_s, drain_s, next_calls

    def column(self, name):
        return self.columns[self.names.index(name)]

    def null(self, name):
        return self.nulls[self.names.index(name)]


class PlanLib(object):
    def __init__(self, path):
        if not os.path.exists(path):
            raise OSError("plan driver library missing: %s" % path)
        self.path = path
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L = self.lib
        L.ssplan_run.restype = C.c_int
        L.ssplan_run.argtypes = [C.c_char_p, C.c_int32, C.POINTER(_Table), C.c_int64, C.c_int32,
                                 C.POINTER(C.c_void_p)]
        for name, res in [("code", C.c_int), ("error", C.c_char_p), ("ncols", C.c_int32),
                          ("rows", C.c_int64), ("create_seconds", C.c_double),
                          ("drain_seconds", C.c_double), ("next_calls", C.c_int64), ("spied_children", C.c_int64)]:
            f = getattr(L, "ssplan_result_" + name)
            f.restype, f.argtypes = res, [C.c_void_p]
        for name, res in [("col_name", C.c_char_p), ("col_dtype", C.c_int32),
                          ("col_nullable", C.c_int32), ("col_data", C.c_void_p),
                          ("col_is_null", C.c_void_p)]:
            f = getattr(L, "ssplan_result_" + name)
            f.restype, f.argtypes = res, [C.c_void_p, C.c_int32]
        L.ssplan_result_col_bytes.restype, L.ssplan_result_col_bytes.argtypes = C.c_void_p, [C.c_void_p, C.c_int32]
        L.ssplan_result_free.restype, L.ssplan_result_free.argtypes = None, [C.c_void_p]
        L.ssplan_impl.restype = C.c_char_p

    @property
    def impl(self):
        return self.lib.ssplan_impl().decode()

    def run(self, plan, tables, next_max_rows=0, flags=0):
        """tables: list of lists of Column. Returns PlanResult (numpy copies)."""
        keep = []
        ctables = (_Table * max(1, len(tables)))()
        for t, cols in enumerate(tables):
            rows = len(cols[0].data) if cols else 0
            carr = (_Column * max(1, len(cols)))()
            for i, c in enumerate(cols):
                assert len(c.data) == rows
                carr[i].name = c.name.encode()
                carr[i].dtype = c.dtype
                carr[i].nullable = 1 if c.nullable else 0
                carr[i].data = c.data.ctypes.data
                carr[i].is_null = None if c.is_null is None else c.is_null.ctypes.data
                keep.append(c)
            keep.append(carr)
            ctables[t].ncols, ctables[t].rows, ctables[t].cols = len(cols), rows, carr
        out = C.c_void_p()
        L = self.lib
        L.ssplan_run(plan.encode(), len(tables), ctables, next_max_rows, flags, C.byref(out))
        try:
            code = L.ssplan_result_code(out)
            error = L.ssplan_result_error(out).decode(errors="replace")
            rows = L.ssplan_result_rows(out)
            n = L.ssplan_result_ncols(out)
            names, dtypes, nullable, columns, nulls = [], [], [], [], []
            for i in range(n):
                names.append(L.ssplan_result_col_name(out, i).decode())
                dt = L.ssplan_result_col_dtype(out, i)
                dtypes.append(dt)
                nullable.append(bool(L.ssplan_result_col_nullable(out, i)))
                if code != 0 or (flags & (SSPLAN_DISCARD | SSPLAN_BIND_ONLY)):
                    columns.append(None)
                    nulls.append(None)
                    continue
                if dt in VARLEN:
                    cells = np.empty(rows, dtype=object)
                    if rows:
                        lens = np.frombuffer((C.c_char * (rows * 8)).from_address(L.ssplan_result_col_data(out, i)),
                                             dtype=np.int64, count=rows)
                        total = int(lens.sum())
                        raw = C.string_at(L.ssplan_result_col_bytes(out, i), total) if total else b""
                        at = 0
                        for k in range(rows):
                            cells[k] = raw[at:at + int(lens[k])]
                            at += int(lens[k])
                    columns.append(cells)
                    q = L.ssplan_result_col_is_null(out, i)
                    if q and rows:
                        nb = (C.c_char * rows).from_address(q)
                        nulls.append(np.frombuffer(nb, dtype=np.uint8, count=rows).astype(np.bool_))
                    else:
                        nulls.append(None)
                    continue
                npdt = np.dtype(NP_OF[dt])
                p = L.ssplan_result_col_data(out, i)
                if rows:
                    buf = (C.c_char * (rows * npdt.itemsize)).from_address(p)
                    columns.append(np.frombuffer(buf, dtype=npdt, count=rows).copy())
                else:
                    columns.append(np.zeros(0, dtype=npdt))
                q = L.ssplan_result_col_is_null(out, i)
                if q and rows:
                    nb = (C.c_char * rows).from_address(q)
                    nulls.append(np.frombuffer(nb, dtype=np.uint8, count=rows).astype(np.bool_))
                else:
                    nulls.append(None)
            res = PlanResult(code, error, names, dtypes, nullable, columns, nulls, rows,
                             L.ssplan_result_create_seconds(out),
                             L.ssplan_result_drain_seconds(out),
                             L.ssplan_result_next_calls(out))
            res.spied_children = L.ssplan_result_spied_children(out)
            return res
        finally:
            L.ssplan_result_free(out)
